// g++ build of the coupled (w, T) line solve (csrc/tfb_joint.h) for the CPU tests.
#include "../../transiflow_b200/csrc/tfb_joint.h"

extern "C" void tfh_joint_lines(int nz, const double* zc, int nmodes, const double* mu, double cv, double cT,
                                double* w, double* T, double* al, double* be) {
    for (int m = 0; m < nmodes; m++) tfb_joint_line(nz, zc, mu[m], cv, cT, nmodes, w + m, T + m, al + m, be + m);
}

// factor once, substitute: must reproduce tfh_joint_lines (fac: 4 arrays of (2 nz - 1) x nmodes doubles)
extern "C" void tfh_joint_factor_substitute(int nz, const double* zc, int nmodes, const double* mu, double cv, double cT,
                                            double* w, double* T, double* fac) {
    const long long span = (long long)(2 * nz - 1) * nmodes;
    for (int m = 0; m < nmodes; m++) {
        tfb_joint_factor_line<double>(nz, zc, mu[m], cv, cT, nmodes, fac + m, fac + span + m, fac + 2 * span + m, fac + 3 * span + m);
        tfb_joint_substitute_line<double, double, double>(nz, zc, cv, cT, nmodes, w + m, T + m, fac + m, fac + span + m,
                                                          fac + 2 * span + m, fac + 3 * span + m);
    }
}
