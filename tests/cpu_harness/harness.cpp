// CPU verification harness: compiles the GENERATED row functions (the exact code the CUDA
// kernels run per thread) with g++ and drives them cell by cell, so their arithmetic can be
// compared bit-for-bit with the oracle in the GPU-less build container.  Test-only.
#include <stdint.h>
#include <string.h>
#include "../../transiflow_b200/csrc/tfb_cell.h"
#include "../../transiflow_b200/csrc/gen/all_configs.h"

struct HostState {
    const TfbGrid* g; const double* s; int i, j, k;
    double operator()(int d, int ox, int oy, int oz) const { return tfb_padded_load(*g, s, 0, i + ox, j + oy, k + oz, d); }
};

template <class Cfg>
static int run(const TfbGrid& g, const TfbParams& prm, const double* state, const double* frc_static,
               int64_t* row_ptr, int64_t* col, double* val, double* rhs, int64_t cap) {
    int64_t idx = 0, row = 0;
    row_ptr[0] = 0;
    for (int k = 0; k < g.nz; k++) for (int j = 0; j < g.ny; j++) for (int i = 0; i < g.nx; i++) {
        TfbCell c;
        tfb_make_cell<Cfg::NFORCE>(g, i, j, k, c);
        HostState P{&g, state, i, j, k};
        for (int d1 = 0; d1 < Cfg::DOF; d1++, row++) {
            double J[32]; double f = 0.0;
            for (int s = 0; s < 32; s++) J[s] = 0.0;
            TfbArraySink sink{J};
            const bool yz_interior = !(c.near[1] || c.far[1] || c.far2[1] ||
                                       (!Cfg::FLAT && (c.near[2] || c.far[2] || c.far2[2])) || (Cfg::ID == 7 && i <= 1 && j <= 1));
            const bool x_interior = !(c.near[0] || c.far[0] || c.far2[0]);
            // the same three instantiations the CUDA kernels use
            if (yz_interior && x_interior) Cfg::template row<true, true, 0>(d1, prm, c, P, sink, f);   // BC-free
            else if (yz_interior) Cfg::template row<true, true, 1>(d1, prm, c, P, sink, f);             // x-face ops only
            else Cfg::template row<true, true, 2>(d1, prm, c, P, sink, f);                               // full recipe
            unsigned m = Cfg::mask(d1, c);
            int ns = Cfg::nslot(d1);
            for (int s = 0; s < ns; s++) if (m >> s & 1u) {
                if (idx >= cap) return -1;
                int d2, dx, dy, dz;
                Cfg::slot(d1, s, d2, dx, dy, dz);
                col[idx] = tfb_column(g, i, j, k, d2, dx, dy, dz);
                val[idx] = J[s];
                idx++;
            } else if (J[s] != 0.0) {
                return -2 - (int)row;   // a masked-out slot must be structurally zero
            }
            rhs[row] = f + (frc_static ? frc_static[row] : 0.0);
            row_ptr[row + 1] = idx;
        }
    }
    return 0;
}

extern "C" int tfh_assemble(int cfg, const TfbGrid* g, const TfbParams* prm, const double* state,
                            const double* frc_static, int64_t* row_ptr, int64_t* col, double* val,
                            double* rhs, int64_t cap) {
#define X(C) if (cfg == C::ID) return run<C>(*g, *prm, state, frc_static, row_ptr, col, val, rhs, cap);
    TFB_FOR_EACH_CONFIG(X)
#undef X
    return -100;
}
