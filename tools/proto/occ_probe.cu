#include <cstdio>
#include "../../transiflow_b200/csrc/tfb_fdm_tc.cuh"
using namespace tfbtc;
int main() {
    auto kern = tfb_fdm_plane_kernel<16, 3>;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    printf("static smem %zu, regs %d, maxThreads %d, maxDyn %d\n", fa.sharedSizeBytes, fa.numRegs, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    for (int kb : {0, 16, 32, 48, 64, 72, 96, 100, 110, 113, 128}) {
        int occ = -1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, (size_t)kb * 1024);
        printf("dyn smem %3d KB -> occ %d (%s)\n", kb, occ, cudaGetErrorString(e));
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("smem per SM %zu, per block optin %zu, reserved %zu, regs/SM %d\n", p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.reservedSharedMemPerBlock, p.regsPerMultiprocessor);
    return 0;
}
