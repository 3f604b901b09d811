// Micro-benchmarks that calibrate the ceilings quoted in DESIGN.md (run on the B200 via gpurun):
//   * copy bandwidth for footprints from L2-resident to HBM-resident
//   * fp64 FMA / ADD issue rate
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n, int reps) {
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            double2 v;
            asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a + i));
            v.x += r;
            b[i] = v;
        }
    }
}
__global__ void k_read(const double2* __restrict__ a, double* out, size_t n, int reps) {
    double s = 0;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            double2 v;
            asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a + i));
            s += v.x + v.y;
        }
    }
    if (s == 1.2345) out[0] = s;
}
template <int ILP, bool ADD>
__global__ void k_fp64(double* out, int iters, double c) {
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = ADD ? __dadd_rn(v[j], c) : fma(v[j], c, 0.5);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j];
    if (s == 1.2345) out[0] = s;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
    size_t maxb = (size_t)1 << 30;
    double2 *a, *b; double* out;
    CK(cudaMalloc(&a, maxb)); CK(cudaMalloc(&b, maxb)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(a, 0, maxb)); CK(cudaMemset(b, 0, maxb));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int nsm = p.multiProcessorCount;
    size_t sizes_mb[] = {4, 8, 16, 24, 32, 48, 64, 96, 256, 1024};
    for (size_t mb : sizes_mb) {
        size_t bytes = mb << 20;               // per buffer
        size_t n = bytes / 16;
        int reps = (int)((((size_t)8) << 30) / bytes); if (reps < 2) reps = 2; if (reps > 400) reps = 400;
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) k_copy<<<nsm * 8, 256>>>(a, b, n, 2); else k_read<<<nsm * 8, 256>>>(a, out, n, 2);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            if (mode == 0) k_copy<<<nsm * 8, 256>>>(a, b, n, reps); else k_read<<<nsm * 8, 256>>>(a, out, n, reps);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double gb = (double)bytes * reps * (mode == 0 ? 2 : 1) / 1e9;
            printf("%s buffer %5zu MB reps %3d : %8.1f GB/s\n", mode == 0 ? "copy(r+w)" : "read     ", mb, reps, gb / (ms * 1e-3));
        }
    }
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int add = 0; add < 2; ++add) {
            const int iters = 20000;
            if (add) k_fp64<8, true><<<nsm, warps * 32>>>(out, 10, 1.0000001); else k_fp64<8, false><<<nsm, warps * 32>>>(out, 10, 1.0000001);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            if (add) k_fp64<8, true><<<nsm, warps * 32>>>(out, iters, 1.0000001); else k_fp64<8, false><<<nsm, warps * 32>>>(out, iters, 1.0000001);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)nsm * warps * 32 * 8.0 * iters;
            printf("fp64 %s warps/SM %2d : %7.2f T instr-lanes/s  (%.1f per clk per SM at 1.965 GHz)\n", add ? "DADD" : "DFMA", warps,
                   ops / (ms * 1e-3) / 1e12, ops / (ms * 1e-3) / nsm / 1.965e9);
        }
    }
    return 0;
}
