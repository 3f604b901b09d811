// Write-pattern micro-benchmark: how fast can B200 absorb the column kernel's store pattern?
// Each warp-level store instruction writes `rows` chunks of `32/rows` lanes x 16 B; chunks of one
// instruction are `row_stride` bytes apart; consecutive warps write adjacent chunks.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// buffer viewed as [n_blocks][64 rows][4096 cols] of double2 (a "transform" = 4 MB); warp w of the grid
// loops over transforms; inside a transform it owns `cw` columns and writes all 64 rows, 8 per instruction
template <int CW>
__global__ void k_pattern(double2* buf, long n_transforms) {
    const int lane = threadIdx.x & 31;
    const int c = lane % CW, bb = lane / CW;   // bb < 32/CW rows per instruction
    constexpr int RPI = 32 / CW;               // rows per instruction
    const long nw = (long)gridDim.x * (blockDim.x >> 5);
    const long wg = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long groups = 4096 / CW;
    const long units = groups * n_transforms;
    for (long u = wg; u < units; u += nw) {
        const long t = u / groups, g = u % groups;
        double2* base = buf + t * (64L * 4096) + g * CW + c;
        const double2 v = make_double2((double)u, (double)lane);
#pragma unroll
        for (int j = 0; j < 64 / RPI; ++j) base[(long)(bb + RPI * j) * 4096] = v;
    }
}
__global__ void k_write(double2* buf, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_double2((double)i, 1.0);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    size_t bytes = (size_t)4 << 30;
    double2* a; CK(cudaMalloc(&a, bytes)); CK(cudaMemset(a, 0, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_write<<<nsm * 8, 256>>>(a, bytes / 16); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        printf("coalesced write 4 GB: %.1f GB/s\n", bytes / 1e9 / (ms * 1e-3));
    }
    const long nt = bytes / (64L * 4096 * 16);
#define RUN(CW) for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_pattern<CW><<<nsm * 4, 128>>>(a, nt); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); \
        cudaEventElapsedTime(&ms, e0, e1); printf("pattern CW=%2d (%3d B chunks, %d rows/instr): %.1f GB/s\n", CW, CW * 16, 32 / CW, bytes / 1e9 / (ms * 1e-3)); }
    RUN(4) RUN(8) RUN(16) RUN(32)
    return 0;
}
