// g++ build of the coupled (w, T) line solve (csrc/tfb_joint.h) for the CPU tests.
#include "../../transiflow_b200/csrc/tfb_joint.h"

extern "C" void tfh_joint_lines(int nz, const double* zc, int nmodes, const double* mu, double cv, double cT,
                                double* w, double* T, double* al, double* be) {
    for (int m = 0; m < nmodes; m++) tfb_joint_line(nz, zc, mu[m], cv, cT, nmodes, w + m, T + m, al + m, be + m);
}
