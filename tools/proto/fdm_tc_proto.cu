// Stand-alone check of csrc/tfb_fdm_tc.cuh (tcgen05 3xTF32 plane transforms) against an fp64 CPU reference,
// and its timing on the 128^3 velocity-solve shape.  Build: see tools/proto/build_proto.sh; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../../transiflow_b200/csrc/tfb_fdm_tc.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

using namespace tfbtc;

template <int KC, int STAGES>
static int run_case(int rows, int cols, int m1, int m2, int narr, int nplanes, bool timing) {
    using G = Geo<KC>;
    std::mt19937_64 rng(1234 + rows * 7 + cols);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    const int k1pad = (cols + KC - 1) / KC * KC, k2pad = (rows + KC - 1) / KC * KC;
    const int n1 = (rows + 15) / 16 * 16, n2 = (cols + 15) / 16 * 16;
    // A1: m1 x cols (rows >= m1 zero), A2: m2 x rows
    std::vector<std::vector<double>> A1(narr), A2(narr);
    std::vector<float*> dA1(narr), dA2(narr);
    for (int q = 0; q < narr; q++) {
        A1[q].resize((size_t)m1 * cols); A2[q].resize((size_t)m2 * rows);
        for (auto& v : A1[q]) v = U(rng) / sqrt((double)cols);
        for (auto& v : A2[q]) v = U(rng) / sqrt((double)rows);
        std::vector<float> f;
        tfb_tc_format_matrix<KC>(A1[q].data(), m1, cols, cols, k1pad, f);
        CK(cudaMalloc(&dA1[q], f.size() * 4)); CK(cudaMemcpy(dA1[q], f.data(), f.size() * 4, cudaMemcpyHostToDevice));
        tfb_tc_format_matrix<KC>(A2[q].data(), m2, rows, rows, k2pad, f);
        CK(cudaMalloc(&dA2[q], f.size() * 4)); CK(cudaMemcpy(dA2[q], f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    }
    const long long plane = (long long)rows * cols;
    std::vector<float> in((size_t)narr * nplanes * plane);
    for (auto& v : in) v = (float)U(rng);
    float *din, *dout;
    CK(cudaMalloc(&din, in.size() * 4)); CK(cudaMalloc(&dout, in.size() * 4));
    CK(cudaMemcpy(din, in.data(), in.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xff, in.size() * 4));
    PlaneArgs a{};
    for (int q = 0; q < narr; q++) {
        a.in[q] = din + (size_t)q * nplanes * plane; a.out[q] = dout + (size_t)q * nplanes * plane;
        a.A1[q] = dA1[q]; a.A2[q] = dA2[q];
    }
    a.narr = narr; a.nplanes = nplanes; a.rows = rows; a.cols = cols; a.plane_stride = plane;
    a.k1pad = k1pad; a.k2pad = k2pad; a.n1 = n1; a.n2 = n2;
    auto kern = tfb_fdm_plane_kernel<KC, STAGES>;
    const size_t smem = plane_kernel_smem<KC, STAGES>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    const long long items = (long long)narr * nplanes;
    const int per_sm = getenv("TFB_PROTO_PER_SM") ? atoi(getenv("TFB_PROTO_PER_SM")) : std::max(std::min(occ, 2), 1);
    const int grid = (int)std::min<long long>(items, 148ll * per_sm);
    kern<<<grid, THREADS, smem>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(in.size());
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    // reference on a few planes
    double maxerr = 0, maxref = 0;
    const int check_items[4] = {0, (int)items / 3, (int)(2 * items / 3), (int)items - 1};
    for (int ci = 0; ci < 4; ci++) {
        const int item = check_items[ci], q = item / nplanes;
        const float* R = in.data() + (size_t)item * plane;
        const float* O = out.data() + (size_t)item * plane;
        std::vector<double> D1((size_t)cols * rows, 0.0);   // D1[a][j]
        for (int aa = 0; aa < m1; aa++)
            for (int j = 0; j < rows; j++) {
                double s = 0;
                for (int i = 0; i < cols; i++) s += A1[q][(size_t)aa * cols + i] * (double)R[(size_t)j * cols + i];
                D1[(size_t)aa * rows + j] = s;
            }
        for (int b = 0; b < rows; b++)
            for (int aa = 0; aa < cols; aa++) {
                double s = 0;
                if (b < m2 && aa < m1)
                    for (int j = 0; j < rows; j++) s += A2[q][(size_t)b * rows + j] * D1[(size_t)aa * rows + j];
                const double got = O[(size_t)b * cols + aa];
                maxerr = std::max(maxerr, fabs(got - s));
                maxref = std::max(maxref, fabs(s));
            }
    }
    printf("KC=%d S=%d rows=%d cols=%d m1=%d m2=%d narr=%d planes=%d occ=%d grid=%d : max abs err %.3e (max ref %.3e, rel %.3e)\n",
           KC, STAGES, rows, cols, m1, m2, narr, nplanes, occ, grid, maxerr, maxref, maxerr / maxref);
    if (timing) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int w = 0; w < 3; w++) kern<<<grid, THREADS, smem>>>(a);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int r = 0; r < reps; r++) kern<<<grid, THREADS, smem>>>(a);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = 1e3 * ms / reps;
        const double flops = 3.0 * 2.0 * items * (2.0 * 128 * (double)n1 * k1pad / 2 + 2.0 * 128 * (double)n2 * k2pad / 2);
        printf("   time %.1f us per launch (%lld items), tensor work %.1f TFLOP/s (3 passes), data %.1f GB/s\n", us, items,
               flops / us / 1e6, 2.0 * in.size() * 4 / us / 1e3);
    }
    cudaFree(din); cudaFree(dout);
    for (int q = 0; q < narr; q++) { cudaFree(dA1[q]); cudaFree(dA2[q]); }
    return maxerr / maxref < 1e-5 ? 0 : 2;
}

int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    int bad = 0;
    if (which == 0 || which == 1) {
        bad |= run_case<32, 2>(128, 128, 128, 128, 1, 4, false);
        bad |= run_case<32, 2>(128, 128, 127, 128, 3, 128, true);
        bad |= run_case<32, 2>(48, 40, 39, 48, 2, 5, false);
        bad |= run_case<32, 2>(20, 12, 12, 19, 1, 3, false);
        bad |= run_case<32, 3>(128, 128, 128, 128, 3, 128, true);
    }
    if (which == 0 || which == 2) {
        bad |= run_case<16, 3>(128, 128, 128, 128, 1, 4, false);
        bad |= run_case<16, 3>(128, 128, 127, 128, 3, 128, true);
        bad |= run_case<16, 3>(48, 40, 39, 48, 2, 5, false);
        bad |= run_case<16, 3>(20, 12, 12, 19, 1, 3, false);
        bad |= run_case<16, 2>(128, 128, 128, 128, 3, 128, true);
        bad |= run_case<16, 4>(128, 128, 128, 128, 3, 128, true);
    }
    printf(bad ? "PROTO FAILED (%d)\n" : "PROTO OK\n", bad);
    return bad;
}
