// dynpy_b200 — dipolar column pass, one THREAD per column (sm_100a).
//
// Two kernels per batch of pair-pairs ("items"):
//
//   k_dd_geom     distances.pdist_ortho (minimum image, distances.py:190-234) + the geometric part of
//                 ddrax.cart_to_dipolar_from_df (ddrax.py:110-126): per pair-frame the unit vector
//                 (x, y, z)/r and r^-3, written once as two double2 streams Gs[pair][2][N]: (x/r, y/r), (z/r, r^-3)
//                 (32 B per pair-frame), plus the per-pair sum of r^-3 (for the mean of G_r).
//                 Pure streaming: 48 B read, 32 B written per pair-frame.
//
//   k_cols_tpc    first pass of the zero-padded forward transform of all 11 series of an item.
//                 M = M1 x 4096, sample n = 4096 n1 + n2.  A thread owns column n2 of one series:
//                 it forms the series' H = M1/2 non-zero samples from the geometry streams
//                 (Y2m in Cartesian closed form), and computes the M1-point column transform as a
//                 radix-2 DIF split into two H-point transforms done entirely in registers by the
//                 generated codelets (even k1: DFT_H(x), odd k1: DFT_H(x W_M1^n1); H = 25 or 32, 16, 8).
//                 No shared memory, no barrier, no exchange between threads: consecutive threads are
//                 consecutive columns, so every load and every 16-byte store of a warp is one
//                 contiguous 256 / 512 byte run.  The inter-pass twiddles W_M^{n2 k1} run as one
//                 complex-product chain per parity from two per-thread seeds.
//
// Per output point: (cost of the two codelets)/M1 + 8 (twiddle chain + product) + ~3 (series
// formation) fp64 instructions = 26 (M1 = 50) / 24 (M1 = 64); 16 B stored, ~12 B loaded.
#pragma once
#include "codelets.cuh"
#include "dd_device.cuh"
#include "fft_engine.cuh"
#include "fft_rows2.cuh"  // DYNPY_T_RS: row stride of T

namespace dynpy {

struct GeomParams {
    const double* X;   // wrapped coordinates [atom][3][N]
    const int* pi;
    const int* pj;
    i64 pair0;         // first pair of the batch
    i64 n_pairs;       // total number of pairs
    i64 n_slots;       // pair slots of the batch (2 * items; slots beyond n_pairs are zero-filled)
    i64 N;
    double a, b, c;
    cd* Gs;            // [n_slots][2][N] double2 streams: (x/r, y/r) and (z/r, r^-3)
    double* rsum;      // [n_slots] sums of r^-3 (zeroed by the caller)
    int pair_fastest;  // 1: grid = (pair slots, frame chunks) — all pairs of the batch sweep one chunk of frames before the
                       // next, so the coordinates of that chunk (A x 3 x 2048 frames) are read from DRAM once per batch
                       // and then from L2; 0: grid = (frame chunks, pair slots), every pair streams its atoms from DRAM
                       // (ncu at cfg2, 1640 pairs per launch: DRAM reads 3.88 -> 1.53 GB, DRAM busy 80 -> 61 %, time 1.36 ->
                       // 1.34 ms: with the reads gone the kernel is bound by load latency at 16 warps per SM, its 2 x 4
                       // frames per thread need the 128 registers — 80 or 64 registers spill 250 - 400 B)
};

// grid: (frame chunks of 1024, pair slots); 256 threads x 4 frames
__global__ void __launch_bounds__(256) k_dd_geom(const GeomParams p) {
    __shared__ double sred[8];
    const i64 slot = p.pair_fastest ? blockIdx.x : blockIdx.y;
    const i64 chunk = p.pair_fastest ? blockIdx.y : blockIdx.x;
    const i64 pr = p.pair0 + slot;
    const bool exists = pr < p.n_pairs;
    const double* xi = p.X;
    const double* xj = p.X;
    if (exists) {
        xi += (i64)__ldg(p.pi + pr) * 3 * p.N;
        xj += (i64)__ldg(p.pj + pr) * 3 * p.N;
    }
    cd* out = p.Gs + slot * 2 * p.N;
    double rs = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const i64 f = chunk * 1024 + e * 256 + threadIdx.x;
        if (f < p.N) {
            double gx = 0.0, gy = 0.0, gz = 0.0, g3 = 0.0;
            if (exists) {
                double dx, dy, dz, sx, sy, sz;
                axis_min(__ldg(xi + f), __ldg(xj + f), p.a, dx, sx);
                axis_min(__ldg(xi + p.N + f), __ldg(xj + p.N + f), p.b, dy, sy);
                axis_min(__ldg(xi + 2 * p.N + f), __ldg(xj + 2 * p.N + f), p.c, dz, sz);
                const double r2 = __dadd_rn(__dadd_rn(sx, sy), sz);
                double rinv = rsqrt(r2);
                rinv = rinv * fma(-0.5 * r2, rinv * rinv, 1.5);  // one Newton step: r^-1 to ~1 ulp
                gx = dx * rinv; gy = dy * rinv; gz = dz * rinv;
                g3 = rinv * rinv * rinv;
            }
            out[f] = make_double2(gx, gy);
            out[p.N + f] = make_double2(gz, g3);
            rs += g3;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rs += __shfl_down_sync(0xffffffffu, rs, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = rs;
    __syncthreads();
    if (threadIdx.x == 0 && exists) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sred[w];
        atomicAdd(p.rsum + slot, t);
    }
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256): four consecutive frames of one coordinate stream per load,
// two output points per store
__device__ __forceinline__ void ldg256(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void stg256(cd* p, cd a, cd b) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// The same kernel for N % 4 == 0 (every stream of every atom then starts on a 32-byte boundary): a thread owns
// 2 x 4 consecutive frames; its 12 loads of 32 bytes (6 coordinate streams x 2) are all issued before the first
// use, so a warp keeps 12 KB in flight instead of 4 x 6 scalar loads of 8 bytes with the arithmetic between
// them; the two output streams are written with 32-byte stores.
// grid: (frame chunks of 2048, pair slots); 256 threads x 2 x 4 frames
__global__ void __launch_bounds__(256) k_dd_geom4(const GeomParams p) {
    __shared__ double sred[8];
    const i64 slot = p.pair_fastest ? blockIdx.x : blockIdx.y;
    const i64 chunk = p.pair_fastest ? blockIdx.y : blockIdx.x;
    const i64 pr = p.pair0 + slot;
    const bool exists = pr < p.n_pairs;
    const double* xi = p.X;
    const double* xj = p.X;
    if (exists) {
        xi += (i64)__ldg(p.pi + pr) * 3 * p.N;
        xj += (i64)__ldg(p.pj + pr) * 3 * p.N;
    }
    cd* out = p.Gs + slot * 2 * p.N;
    const double L[3] = {p.a, p.b, p.c};
    double rs = 0.0;
    double ci[2][3][4], cj[2][3][4];
    i64 f0[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        f0[e] = chunk * 2048 + e * 1024 + threadIdx.x * 4;
        if (f0[e] < p.N && exists) {
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                ldg256(xi + ax * p.N + f0[e], ci[e][ax]);
                ldg256(xj + ax * p.N + f0[e], cj[e][ax]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        if (f0[e] < p.N) {
            cd g0[4], g1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double gx = 0.0, gy = 0.0, gz = 0.0, g3 = 0.0;
                if (exists) {
                    double d[3], s2[3];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) axis_min(ci[e][ax][q], cj[e][ax][q], L[ax], d[ax], s2[ax]);
                    const double r2 = __dadd_rn(__dadd_rn(s2[0], s2[1]), s2[2]);
                    double rinv = rsqrt(r2);
                    rinv = rinv * fma(-0.5 * r2, rinv * rinv, 1.5);  // one Newton step: r^-1 to ~1 ulp
                    gx = d[0] * rinv; gy = d[1] * rinv; gz = d[2] * rinv;
                    g3 = rinv * rinv * rinv;
                }
                g0[q] = make_double2(gx, gy);
                g1[q] = make_double2(gz, g3);
                rs += g3;
            }
            stg256(out + f0[e], g0[0], g0[1]);
            stg256(out + f0[e] + 2, g0[2], g0[3]);
            stg256(out + p.N + f0[e], g1[0], g1[1]);
            stg256(out + p.N + f0[e] + 2, g1[2], g1[3]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rs += __shfl_down_sync(0xffffffffu, rs, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = rs;
    __syncthreads();
    if (threadIdx.x == 0 && exists) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sred[w];
        atomicAdd(p.rsum + slot, t);
    }
}

// ---------------------------------------------------------------- TMA-staged form of the geometry kernel
// The six coordinate streams of a pair are moved by the TMA engine (cp.async.bulk global -> shared, completion
// counted on an mbarrier) through a ring of GEOM_STAGES tiles of GEOM_TILE frames; no thread issues a load for
// them, the LSU only sees the conflict-free shared-memory reads and the 32-byte stores of the two output streams.
// A CTA walks a contiguous range of (pair slot, frame tile) units, so its ring never drains between pairs.
// Needs N % 2 == 0 (bulk copies move multiples of 16 bytes from 16-byte aligned addresses).
#define GEOM_TILE 512
#define GEOM_STAGES 4
#define GEOM_NT 256
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(GEOM_NT, 2) k_dd_geom_tma(const GeomParams p) {
    extern __shared__ __align__(128) unsigned char geom_smem[];
    double* buf = reinterpret_cast<double*>(geom_smem);                       // [STAGES][6][TILE]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(buf + GEOM_STAGES * 6 * GEOM_TILE);
    __shared__ double sred[GEOM_NT / 32];
    const i64 tiles = (p.N + GEOM_TILE - 1) / GEOM_TILE;
    const i64 units = p.n_slots * tiles;
    const i64 u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    if (u0 >= u1) return;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < GEOM_STAGES; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // producer (thread 0): the six streams of unit u into stage s
    auto issue = [&](i64 u, int s) {
        const i64 slot = u / tiles, tile = u - slot * tiles;
        const i64 pr = p.pair0 + slot;
        if (pr >= p.n_pairs) {   // slot beyond the last pair: nothing to copy, complete the phase by a plain arrive
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + s)) : "memory");
            return;
        }
        const i64 f0 = tile * GEOM_TILE;
        const unsigned bytes = (unsigned)((p.N - f0 < GEOM_TILE ? p.N - f0 : GEOM_TILE) * 8);
        const double* xi = p.X + (i64)__ldg(p.pi + pr) * 3 * p.N + f0;
        const double* xj = p.X + (i64)__ldg(p.pj + pr) * 3 * p.N + f0;
        double* dst = buf + (i64)s * 6 * GEOM_TILE;
        mbar_expect_tx(full + s, 6 * bytes);
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            tma_load_1d(dst + (2 * ax) * GEOM_TILE, xi + ax * p.N, bytes, full + s);
            tma_load_1d(dst + (2 * ax + 1) * GEOM_TILE, xj + ax * p.N, bytes, full + s);
        }
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < GEOM_STAGES; ++s)
            if (u0 + s < u1) issue(u0 + s, s);
    const double L[3] = {p.a, p.b, p.c};
    double rs = 0.0;
    i64 cur_slot = u0 / tiles;
    auto flush = [&](i64 slot) {
        double t = rs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0 && p.pair0 + slot < p.n_pairs) {
            double a = 0.0;
#pragma unroll
            for (int w = 0; w < GEOM_NT / 32; ++w) a += sred[w];
            atomicAdd(p.rsum + slot, a);
        }
        __syncthreads();
        rs = 0.0;
    };
    for (i64 u = u0; u < u1; ++u) {
        const int s = (int)((u - u0) % GEOM_STAGES);
        const unsigned parity = (unsigned)(((u - u0) / GEOM_STAGES) & 1);
        const i64 slot = u / tiles, tile = u - slot * tiles;
        if (slot != cur_slot) { flush(cur_slot); cur_slot = slot; }
        const bool exists = p.pair0 + slot < p.n_pairs;
        mbar_wait(full + s, parity);
        const double2* b = reinterpret_cast<const double2*>(buf + (i64)s * 6 * GEOM_TILE);
        const i64 f = tile * GEOM_TILE + 2 * threadIdx.x;     // two consecutive frames per thread
        double c6[6][2];                                       // 16-byte shared loads: a quarter warp reads 128 contiguous bytes
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double2 t = b[k * (GEOM_TILE / 2) + threadIdx.x];
            c6[k][0] = t.x; c6[k][1] = t.y;
        }
        cd g0[2], g1[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double gx = 0.0, gy = 0.0, gz = 0.0, g3 = 0.0;
            if (exists && f + q < p.N) {
                double d[3], s2[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) axis_min(c6[2 * ax][q], c6[2 * ax + 1][q], L[ax], d[ax], s2[ax]);
                const double r2 = __dadd_rn(__dadd_rn(s2[0], s2[1]), s2[2]);
                double rinv = rsqrt(r2);
                rinv = rinv * fma(-0.5 * r2, rinv * rinv, 1.5);  // one Newton step: r^-1 to ~1 ulp
                gx = d[0] * rinv; gy = d[1] * rinv; gz = d[2] * rinv;
                g3 = rinv * rinv * rinv;
            }
            g0[q] = make_double2(gx, gy);
            g1[q] = make_double2(gz, g3);
            rs += g3;
        }
        if (f < p.N) {   // N is even: frames f and f + 1 exist together
            cd* out = p.Gs + slot * 2 * p.N;
            stg256(out + f, g0[0], g0[1]);
            stg256(out + p.N + f, g1[0], g1[1]);
        }
        __syncthreads();                                        // every thread has read stage s: refill it
        if (threadIdx.x == 0 && u + GEOM_STAGES < u1) issue(u + GEOM_STAGES, s);
    }
    flush(cur_slot);
}

struct ColsTpcParams {
    const cd* Gs;      // [2*items][2][N] geometry streams of the batch
    i64 items;
    i64 pair0, n_pairs; // first pair of the batch / total pairs (pair q of the last item may not exist)
    i64 N;
    cd* T;             // [items][11][M1][DYNPY_T_RS]
    i64 M;
    const cd* seeds;   // [4096][2]: W_M^{n2}, W_M^{2 n2}
    int sync_units;    // 1: CTA barrier once per series
    const double* rsum; // [2*items] sums of r^-3 from k_dd_geom (complete before this kernel starts)
    double inv_n;
};

// twiddle seeds of every column (filled once per call)
__global__ void k_fill_seeds(cd* seeds, double inv_M) {
    const int n2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (n2 < 4096) {
        seeds[2 * n2] = w_exact(n2, inv_M);
        seeds[2 * n2 + 1] = w_exact(2 * (i64)n2, inv_M);
    }
}

// 16-byte streaming store: T is written once and read once by the row kernel, it must not push the
// geometry streams (re-read by the 11 series of an item) out of L2
__device__ __forceinline__ void st_stream(cd* p, cd v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// threads per CTA of the column kernel: 256 x 2 CTAs per SM (8 warps running the same straight-line code
// close together keep the instruction caches effective: 128 x 4 lost 12 %; two independent CTAs overlap
// their load and compute phases: 512 x 1 lost 5 %; 128 x 2-3 or 256 x 1 at 168-255 registers: 8 % slower for
// M1 = 50); 128 x 2 CTAs at 255 registers for M1 = 64, whose 32-point halves spill at anything less
#ifndef TPC_NT
#define TPC_NT 256
#endif
#ifndef TPC_PER_SM
#define TPC_PER_SM (512 / TPC_NT)
#endif
// 1 (A/B, with TPC_NT = 128 and TPC_PER_SM = 2): the zone holds BOTH streams of a series (2 H entries per thread), the
// second stream of the next series is copied under the even half, the first under the odd half: no copy is waited for
#ifndef TPC_DUAL
#define TPC_DUAL 0
#endif
#ifndef TPC64_PER_SM
#define TPC64_PER_SM 2   // M1 = 64: 32-point halves need ~250 registers; 2 x 128 threads without spills beat 3 x 128 with
#endif
template <int M1> struct TpcShape {
    static constexpr int NT = M1 >= 64 ? 128 : TPC_NT;
    static constexpr int PER_SM = M1 >= 64 ? TPC64_PER_SM : TPC_PER_SM;
};

template <int H> struct ColCodelets;
template <> struct ColCodelets<25> {
    static __device__ __forceinline__ void even(cd (&v)[25]) { fft25(v); }
    static __device__ __forceinline__ void odd(cd (&v)[25]) { fft25_odd(v); }
};
template <> struct ColCodelets<10> {
    static __device__ __forceinline__ void even(cd (&v)[10]) { fft10(v); }
    static __device__ __forceinline__ void odd(cd (&v)[10]) { fft10_odd(v); }
};
template <> struct ColCodelets<12> {
    static __device__ __forceinline__ void even(cd (&v)[12]) { fft12(v); }
    static __device__ __forceinline__ void odd(cd (&v)[12]) { fft12_odd(v); }
};
template <> struct ColCodelets<20> {
    static __device__ __forceinline__ void even(cd (&v)[20]) { fft20(v); }
    static __device__ __forceinline__ void odd(cd (&v)[20]) { fft20_odd(v); }
};
template <> struct ColCodelets<32> {
    static __device__ __forceinline__ void even(cd (&v)[32]) { fft32(v); }
    static __device__ __forceinline__ void odd(cd (&v)[32]) { fft32_odd(v); }
};
template <> struct ColCodelets<16> {
    static __device__ __forceinline__ void even(cd (&v)[16]) { fft16(v); }
    static __device__ __forceinline__ void odd(cd (&v)[16]) { fft16_odd(v); }
};
template <> struct ColCodelets<8> {
    static __device__ __forceinline__ void even(cd (&v)[8]) { fft8(v); }
    static __device__ __forceinline__ void odd(cd (&v)[8]) { fft8_odd(v); }
};

// Work units u = (item, column block of 128, slot), slot fastest, dealt round-robin to the CTAs: the 11
// series of an (item, column block) run on 11 CTAs at about the same time and share the geometry
// streams through L2.  Series formation has two branch-free families (keeps the kernel inside the
// instruction cache):  complex  s (x + i y)^e  with s = c [r^-3] [z]  (Y21, Y22, F21, F22), and the
// real series of the two pairs of an item packed as re + i im  (Y20, r^-3, F20).
// 16-byte global -> shared copy that zero-fills when `on` is false
__device__ __forceinline__ void cp_async16_zfill(cd* smem_dst, const cd* gsrc, bool on) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = on ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

// Immediate-offset forms: the H copies of a stream (and the M1 stores of a series) differ from one base address by
// compile-time constants, which fit the 24-bit offset field of LDGSTS / STG.  With the addresses passed as computed
// pointers the compiler materialised every 64-bit address (and every source size) before the first copy: 75
// registers per burst, spilled.
template <int SOFF, int GOFF>
__device__ __forceinline__ void cp_async16_zfill_at(unsigned sa, const cd* gsrc, bool off) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %2, 0;\n"
        "cp.async.cg.shared.global [%0 + %3], [%1 + %4], 16, p;\n"   // p: ignore-src, the 16 bytes are zero-filled
        "}\n" ::"r"(sa), "l"(gsrc), "r"((int)off), "n"(SOFF), "n"(GOFF) : "memory");
}
template <int STRIDE_S, int STRIDE_G, int... J>
__device__ __forceinline__ void cp_async_burst(unsigned sa, const cd* gsrc, int jlim, std::integer_sequence<int, J...>) {
    (cp_async16_zfill_at<J * STRIDE_S, J * STRIDE_G>(sa, gsrc, J >= jlim), ...);
}
template <i64 GOFF>
__device__ __forceinline__ void st_stream_at(cd* p, cd v) {
    asm volatile("st.global.cs.v2.f64 [%0 + %3], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y), "n"(GOFF) : "memory");
}

// outputs k1 = 2 a + PAR of a column, a = A..H-1, times the running inter-pass twiddle (w advances by w2 per output;
// the first even output carries no twiddle)
template <int H, int PAR, int A>
__device__ __forceinline__ void tpc_store_chain(cd* dst, const cd (&v)[H], cd w, const cd w2, bool live) {
    if constexpr (A < H) {
        constexpr bool plain = PAR == 0 && A == 0;
        const cd val = plain ? v[A] : cmul(v[A], w);
        if (live) st_stream_at<(i64)(2 * A + PAR) * DYNPY_T_RS * (i64)sizeof(cd)>(dst, val);
        if (!plain && A + 1 < H) w = cmul(w, w2);
        tpc_store_chain<H, PAR, A + 1>(dst, v, w, w2, live);
    }
}

template <int K> struct TpcKind { static constexpr int value = K; };

template <int M1>
__global__ void __launch_bounds__(TpcShape<M1>::NT, TpcShape<M1>::PER_SM) k_cols_tpc(const ColsTpcParams p) {
    constexpr int H = M1 / 2;
    constexpr int TPC_THREADS = TpcShape<M1>::NT;
    constexpr int NCB = (4096 + TPC_THREADS - 1) / TPC_THREADS;  // the last column block may be partial
    constexpr bool RAGGED = (4096 % TPC_THREADS) != 0;
    constexpr bool DUAL = TPC_DUAL != 0 && M1 < 64;
    extern __shared__ double2 smem[];
    // thread-private zone of H double2 at zone[j * TPC_THREADS]: landing area of the two geometry streams (one
    // cp.async burst each: all H loads of a stream are in flight together, L2 latency is paid twice per
    // series instead of once per frame), then the stash of the H samples for the odd half
    cd* zone = smem + threadIdx.x;
    constexpr int nslots = 11;
    const i64 units = p.items * NCB * nslots;
    const int nfr = (int)(p.N < (i64)H * 4096 ? p.N : (i64)H * 4096);  // frames beyond M/2 do not exist
    constexpr double C20 = 0.31539156525252000603, C21 = -0.77254840404638362, C22 = 0.38627420202319181;
    constexpr double KF = 1.5853309190424043;
    constexpr i64 TS = (i64)M1 * DYNPY_T_RS;
    const i64 N = p.N;

    // stream A of a unit: (x/r, y/r) of the series' pair, or (z/r, r^-3) of pair p for the packed series
    auto issue_stream = [&](i64 u, int which) {
        if (u < units) {
            const int slot = (int)(u % nslots);
            const i64 ic = u / nslots;
            const int n2 = (int)(ic % NCB) * TPC_THREADS + threadIdx.x;
            const i64 item = ic / NCB;
            const bool packed = slot == 0 || slot == 5 || slot == 6;
            const bool second = slot == 2 || slot == 4 || slot == 8 || slot == 10;
            // packed: A = (z, r3) of p, B = (z, r3) of q;  complex: A = (x, y), B = (z, r3) of the series' pair
            const i64 sl = packed ? (which ? 3 : 1) : (second ? 2 : 0) + which;
            const cd* src = p.Gs + item * 4 * N + sl * N + n2;
            const int jlim = (nfr > n2 && (!RAGGED || n2 < 4096)) ? (nfr - n2 + 4095) >> 12 : 0;
            // rows j >= jlim are zero-filled (src-size 0: nothing is read; the address still lies inside the
            // workspace, whose T region follows the geometry streams)
            cd* dstz = (DUAL && which) ? zone + H * TPC_THREADS : zone;
            if (!(DUAL && which && (slot == 3 || slot == 4)))   // Y22 has no second stream
                cp_async_burst<TPC_THREADS * 16, 4096 * 16>((unsigned)__cvta_generic_to_shared(dstz), src, jlim,
                                                             std::make_integer_sequence<int, H>{});
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const cd* zoneB = DUAL ? zone + H * TPC_THREADS : zone;   // where the second stream of a series is read

    issue_stream(blockIdx.x, 0);
    if (DUAL) issue_stream(blockIdx.x, 1);
    for (i64 u = blockIdx.x; u < units; u += gridDim.x) {
        const int slot = (int)(u % nslots);
        const i64 ic = u / nslots;
        const int n2 = (int)(ic % NCB) * TPC_THREADS + threadIdx.x;
        const i64 item = ic / NCB;
        const bool live = !RAGGED || n2 < 4096;   // threads past the last column run on zeros and store nothing
        const cd w1 = ldg_cd(p.seeds + 2 * (n2 & 4095)), w2 = ldg_cd(p.seeds + 2 * (n2 & 4095) + 1);
        cd* dst = p.T + (T_ITEM(item) * nslots + slot) * TS + n2;
        // slot -> series (kernels.h): 0 Y20 p+iq | 1,2 Y21 p,q | 3,4 Y22 p,q | 5 r^-3 p+iq | 6 F20 p+iq | 7,8 F21 p,q | 9,10 F22 p,q
        const int jlim = (nfr > n2 && live) ? (nfr - n2 + 4095) >> 12 : 0;
        if (p.sync_units) __syncthreads();  // re-align the warps of the CTA on the same code (instruction-cache locality)
        cd v[H];
        // Series formation, one specialised straight-line block per kind of series (the slot is uniform over the
        // CTA: a branch, no selects).  Frames beyond the series (j >= jlim) and the missing second pair of a last
        // odd item arrive as zero geometry (zero-filled copies / zero-filled streams), which makes every series
        // that carries a factor x, y, z or r^-3 vanish by itself; only Y20 = C20 (3 z^2 - 1) and r^-3 - <r^-3>
        // need their constant term switched off there.
        auto form = [&](auto kind) {
            constexpr int K = decltype(kind)::value;   // 0 Y20 | 5 r^-3 | 6 F20 | 1 Y21 | 3 Y22 | 7 F21 | 9 F22
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if constexpr (K == 0 || K == 5 || K == 6) {
                const bool q_exists = p.pair0 + 2 * item + 1 < p.n_pairs;
                const int jlim_q = q_exists ? jlim : 0;
                const double mean_p = K == 5 ? __ldg(p.rsum + 2 * item) * p.inv_n : 0.0;
                const double mean_q = K == 5 ? __ldg(p.rsum + 2 * item + 1) * p.inv_n : 0.0;   // 0 for a missing pair
                auto real_word = [&](const cd zr, int j, int lim, double mean) -> double {
                    if constexpr (K == 0) return (j < lim ? C20 : 0.0) * fma(3.0 * zr.x, zr.x, -1.0);
                    else if constexpr (K == 5) return zr.y - (j < lim ? mean : 0.0);
                    else return (KF * C20) * zr.y * fma(3.0 * zr.x, zr.x, -1.0);
                };
#pragma unroll
                for (int j = 0; j < H; ++j) v[j].x = real_word(zone[j * TPC_THREADS], j, jlim, mean_p);
                if (!DUAL) {
                    issue_stream(u, 1);
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
#pragma unroll
                for (int j = 0; j < H; ++j) v[j].y = real_word(zoneB[j * TPC_THREADS], j, jlim_q, mean_q);
            } else {
                constexpr bool order1 = K == 1 || K == 7;
                constexpr bool weighted = K >= 7;
                constexpr double c = (order1 ? C21 : C22) * (weighted ? KF : 1.0);
                if constexpr (K == 3) {
                    // Y22 = C22 (x + i y)^2 needs neither z nor r^-3: no second stream (and none of its latency)
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        const cd xy = zone[j * TPC_THREADS];
                        v[j] = make_double2(c * (xy.x * xy.x - xy.y * xy.y), (2.0 * c) * xy.x * xy.y);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        const cd xy = zone[j * TPC_THREADS];
                        v[j] = order1 ? xy : make_double2(xy.x * xy.x - xy.y * xy.y, 2.0 * xy.x * xy.y);
                    }
                    if (!DUAL) {
                        issue_stream(u, 1);
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        const cd zr = zoneB[j * TPC_THREADS];
                        const double sc = c * (weighted ? zr.y : 1.0) * (order1 ? zr.x : 1.0);
                        v[j] = make_double2(sc * v[j].x, sc * v[j].y);
                    }
                }
            }
        };
        switch (slot) {
            case 0: form(TpcKind<0>{}); break;
            case 5: form(TpcKind<5>{}); break;
            case 6: form(TpcKind<6>{}); break;
            case 1: case 2: form(TpcKind<1>{}); break;
            case 3: case 4: form(TpcKind<3>{}); break;
            case 7: case 8: form(TpcKind<7>{}); break;
            default: form(TpcKind<9>{}); break;
        }
#pragma unroll
        for (int j = 0; j < H; ++j) zone[j * TPC_THREADS] = v[j];
        if (DUAL) issue_stream(u + gridDim.x, 1);   // the second half of the zone is free: next series' second stream
        // even outputs
        ColCodelets<H>::even(v);
        tpc_store_chain<H, 0, 0>(dst, v, w2, w2, live);
        // odd outputs
#pragma unroll
        for (int j = 0; j < H; ++j) v[j] = zone[j * TPC_THREADS];
        issue_stream(u + gridDim.x, 0);  // the zone is free: next series' first stream lands under the odd half
        ColCodelets<H>::odd(v);
        tpc_store_chain<H, 1, 0>(dst, v, w1, w2, live);
    }
}

// ---------------------------------------------------------------- long columns: M1 = 16 B (B = 16, 32)
// Two-level form of the same idea for the column lengths of long trajectories (M1 = 512 for 1e6 frames).
// n1 = B a + b (a < 16, only a < 8 is non-zero: the series fills at most half of M), k1 = c + 16 d:
//   X[c + 16 d] = sum_b W_B^{b d} W_M1^{b c} sum_a x[B a + b] W_16^{a c}.
// Phase A, thread (b, column): the 8 samples of its residue b are formed from the geometry streams, the pruned
//   16-point transform over a is the fft8 / fft8_odd pair, the result times W_M1^{b c} goes to shared memory.
// Phase C, thread (c, parity, column): the B-point transform over b as a radix-2 split in registers
//   (fft_{B/2} of y[b] + y[b + B/2], fft_{B/2}_odd of the difference), inter-pass twiddle chain, 128-byte stores.
// 384 threads per SM (3 CTAs of 4 columns: 168 registers, 68 B of spills) beat 512 at 128 registers.
// A CTA keeps its tile of C columns (B C threads) and walks through the series of its items, so the twiddle
// seeds are loop-invariant and the geometry tile it re-reads for the 11 series of an item stays in L2.
template <int B> struct Tpc2Codelets;
template <> struct Tpc2Codelets<32> {
    static __device__ __forceinline__ void even(cd (&v)[16]) { fft16(v); }
    static __device__ __forceinline__ void odd(cd (&v)[16]) { fft16_odd(v); }
};

template <int B, int C>
__global__ void __launch_bounds__(B * C, 384 / (B * C)) k_cols_tpc2(const ColsTpcParams p, const cd* __restrict__ twp) {
    static_assert(B == 32, "phase A (B x C threads) and phase C (16 x 2 x C threads) use the same threads only for B = 32");
    constexpr int M1 = 16 * B;       // C = columns per CTA (B * C threads)
    constexpr int HB = B / 2;
    extern __shared__ double2 smem[];  // S[c][b][col], 16 * B * C entries
    const int tid = threadIdx.x;
    const i64 N = p.N;
    const int nfr = (int)(N < (i64)(M1 / 2) * 4096 ? N : (i64)(M1 / 2) * 4096);
    const double C20 = 0.31539156525252000603, C21 = -0.77254840404638362, C22 = 0.38627420202319181;
    const double KF = 1.5853309190424043;
    constexpr i64 TS = (i64)M1 * DYNPY_T_RS;
    const int tile = blockIdx.x;
    // phase A role
    const int colA = tid % C, bA = tid / C;
    const int n2A = tile * C + colA;
    const cd tb = ldg_cd(twp + bA * (4096 / M1));       // W_M1^b
    // phase C role
    const int colC = tid % C, par = (tid / C) & 1, cC = tid / (2 * C);
    const int n2C = tile * C + colC;
    const double inv_M = 1.0 / (double)p.M;
    const cd seed = w_exact(((i64)n2C * (cC + 16 * par)) % p.M, inv_M);   // W_M^{n2 (c + 16 p)}
    const cd step = w_exact(((i64)n2C * 32) % p.M, inv_M);                // W_M^{32 n2}

    for (i64 item = blockIdx.y; item < p.items; item += gridDim.y) {
        const cd* g = p.Gs + item * 4 * N;
        const bool q_exists = p.pair0 + 2 * item + 1 < p.n_pairs;
        for (int slot = 0; slot < 11; ++slot) {
            const bool packed = slot == 0 || slot == 5 || slot == 6;
            const bool second = slot == 2 || slot == 4 || slot == 8 || slot == 10;
            const bool weighted = slot >= 6;
            const bool order1 = slot == 1 || slot == 2 || slot == 7 || slot == 8;
            // ---- phase A: samples n1 = B a + b, a = 0..7
            cd v[8];
            {
                const cd* sa = g + (packed ? 1 : (second ? 2 : 0)) * N;   // packed: (z, r3) of p ; complex: (x, y)
                const cd* sb = g + (packed ? 3 : (second ? 3 : 1)) * N;   // packed: (z, r3) of q ; complex: (z, r3)
                cd ra[8], rb[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const i64 fr = (i64)(B * a + bA) * 4096 + n2A;
                    const bool on = fr < nfr;
                    ra[a] = on ? ldg_cd(sa + fr) : make_double2(0.0, 0.0);
                    rb[a] = on ? ldg_cd(sb + fr) : make_double2(0.0, 0.0);
                }
                if (packed) {
                    const double c = slot == 5 ? 1.0 : (weighted ? KF * C20 : C20);
                    const double mp = slot == 5 ? __ldg(p.rsum + 2 * item) * p.inv_n : 0.0;
                    const double mq = slot == 5 ? __ldg(p.rsum + 2 * item + 1) * p.inv_n : 0.0;
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const bool on = (i64)(B * a + bA) * 4096 + n2A < nfr;
                        const double yp = slot == 5 ? 1.0 : fma(3.0 * ra[a].x, ra[a].x, -1.0);
                        const double yq = slot == 5 ? 1.0 : fma(3.0 * rb[a].x, rb[a].x, -1.0);
                        v[a].x = on ? (slot == 0 ? c : c * ra[a].y) * yp - mp : 0.0;
                        v[a].y = (on && q_exists) ? (slot == 0 ? c : c * rb[a].y) * yq - mq : 0.0;
                    }
                } else {
                    const double c = (order1 ? C21 : C22) * (weighted ? KF : 1.0);
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const cd xy = ra[a], zr = rb[a];
                        const cd w = order1 ? xy : make_double2(xy.x * xy.x - xy.y * xy.y, 2.0 * xy.x * xy.y);
                        const double sc = c * (weighted ? zr.y : 1.0) * (order1 ? zr.x : 1.0);
                        v[a] = make_double2(sc * w.x, sc * w.y);
                    }
                }
            }
            {
                cd e[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) e[a] = v[a];
                fft8(e);        // y[2 c']
                fft8_odd(v);    // y[2 c' + 1]
                // times W_M1^{b c}: chain over c
                cd w = tb;      // c = 1
                smem[(0 * B + bA) * C + colA] = e[0];
#pragma unroll
                for (int c = 1; c < 16; ++c) {
                    const cd y = (c & 1) ? v[c >> 1] : e[c >> 1];
                    smem[(c * B + bA) * C + colA] = cmul(y, w);
                    if (c < 15) w = cmul(w, tb);
                }
            }
            __syncthreads();
            // ---- phase C: B-point transform over b for (c, column), parity `par`
            cd u[HB];
#pragma unroll
            for (int j = 0; j < HB; ++j) {
                const cd y0 = smem[(cC * B + j) * C + colC];
                const cd y1 = smem[(cC * B + j + HB) * C + colC];
                u[j] = par ? make_double2(y0.x - y1.x, y0.y - y1.y) : make_double2(y0.x + y1.x, y0.y + y1.y);
            }
            __syncthreads();   // the buffer may be rewritten by the next series
            if (par) Tpc2Codelets<B>::odd(u); else Tpc2Codelets<B>::even(u);
            cd* dst = p.T + (item * 11 + slot) * TS + n2C;
            {
                cd w = seed;
#pragma unroll
                for (int dd = 0; dd < HB; ++dd) {
                    const int k1 = cC + 16 * (2 * dd + par);
                    st_stream(dst + (i64)k1 * DYNPY_T_RS, cmul(u[dd], w));
                    if (dd + 1 < HB) w = cmul(w, step);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- long columns, geometry held in registers
// Same two-level transform as k_cols_tpc2, reorganised around what its profile showed (exposed load latency at
// the top of every series, 37 % issue utilisation): the 8 frames a phase-A thread needs are THE SAME for all 11
// series of an item, so the four geometry streams are loaded once per item into registers (32 loads of 16 bytes
// in flight per thread, one exposed latency per item instead of per series, 1/11 of the L2 traffic) and the 11
// series are formed from them.  The phase-A twiddles W_M1^{b c} are loop-invariant per thread: they come from a
// 16 x B table in shared memory (the C column threads of one b read the same entry: one wavefront per warp
// load) instead of a 14-product chain per series.
// B = 16 (M1 = 256) and B = 8 (M1 = 128) run the same two phases with other thread roles: for B = 16 a phase-C thread
// (c, column) does the whole 16-point transform over b (no parity split), 16 C threads in both phases; for B = 8 phase C
// is a plain fft8 per (c, column) on 16 C threads and phase A, which has only 8 C (b, column) roles, is split over two
// threads per role by output parity (one runs fft8 for the even c, the other fft8_odd for the odd c).
template <int B, int C> struct Tpc3Shape {
    static constexpr bool SPLIT_C = B == 32;            // phase C: B-point transform as two B/2-point codelets on two threads
    static constexpr bool SPLIT_A = B == 8;             // phase A: even / odd c on two threads
    static constexpr int NT = (B >= 16 ? B : 16) * C;
    static constexpr int U = SPLIT_C ? B / 2 : B;       // points per phase-C thread
    // row of the exchange buffer S[c][b][col]: B C entries + a pad of 4 (64 bytes) — with C = 4 the four c rows a
    // phase-C warp reads at once are otherwise a multiple of 128 bytes apart (4-way bank conflict on every load)
    static constexpr int SROW = B * C + (C == 4 ? 4 : 0);
    static constexpr size_t smem_bytes(bool gs) {
        return ((size_t)16 * SROW + 16 * B + (gs ? 16 * NT : 0)) * sizeof(cd);
    }
};
// GS = true: the two geometry streams of a series group live in a thread-private shared-memory zone (16 entries per
// thread, filled by cp.async) instead of 64 registers, which lets three CTAs per SM run without spills.
template <int B, int C, int PER_SM, bool GS>
__global__ void __launch_bounds__(Tpc3Shape<B, C>::NT, PER_SM) k_cols_tpc3(const ColsTpcParams p, const cd* __restrict__ twp) {
    static_assert(B == 32 || B == 16 || B == 8, "column lengths 512, 256, 128");
    using Shape = Tpc3Shape<B, C>;
    constexpr int M1 = 16 * B;
    constexpr int HB = Shape::U;
    extern __shared__ double2 smem[];  // S[c][b][col] (16 rows of SROW entries), TA[c][b] (16 * B entries), ZG[2][8][NT]
    constexpr int SROW = Shape::SROW;
    cd* TA = smem + 16 * SROW;
    cd* ZG = TA + 16 * B + threadIdx.x;   // GS: entry (stream s, sample a) at ZG[(8 s + a) * NT]
    const int tid = threadIdx.x;
    const i64 N = p.N;
    const int nfr = (int)(N < (i64)(M1 / 2) * 4096 ? N : (i64)(M1 / 2) * 4096);
    constexpr double C20 = 0.31539156525252000603, C21 = -0.77254840404638362, C22 = 0.38627420202319181;
    constexpr double KF = 1.5853309190424043;
    constexpr i64 TS = (i64)M1 * DYNPY_T_RS;
    const int tile = blockIdx.x;
    const int colA = tid % C, bA = (tid / C) % B, halfA = tid / (B * C);   // halfA: 0 unless SPLIT_A
    const int n2A = tile * C + colA;
    const int colC = tid % C;
    const int par = Shape::SPLIT_C ? (tid / C) & 1 : 0;
    const int cC = Shape::SPLIT_C ? tid / (2 * C) : tid / C;
    const int n2C = tile * C + colC;
    const double inv_M = 1.0 / (double)p.M;
    constexpr int KSTEP = Shape::SPLIT_C ? 32 : 16;                          // k1 = c + 16 par + KSTEP d
    const cd seed = w_exact(((i64)n2C * (cC + 16 * par)) % p.M, inv_M);   // W_M^{n2 (c + 16 p)}
    const cd step = w_exact(((i64)n2C * KSTEP) % p.M, inv_M);             // W_M^{KSTEP n2}
    for (int e = tid; e < 16 * B; e += Shape::NT) TA[e] = ldg_cd(twp + (e / B) * (e % B) * (4096 / M1));   // W_M1^{b c}
    __syncthreads();
    // frames of this thread's residue: (B a + b) 4096 + n2, a = 0..7
    unsigned onmask = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) onmask |= ((i64)(B * a + bA) * 4096 + n2A < nfr) ? (1u << a) : 0u;

    for (i64 item = blockIdx.y; item < p.items; item += gridDim.y) {
        const cd* g = p.Gs + item * 4 * N;
        const bool q_exists = p.pair0 + 2 * item + 1 < p.n_pairs;
        const double mp = __ldg(p.rsum + 2 * item) * p.inv_n;
        const double mq = __ldg(p.rsum + 2 * item + 1) * p.inv_n;
        // series in three groups that share their two streams: complex series of p (streams 0, 1), of q (2, 3),
        // packed real series of p + i q (1, 3)
        cd GA[GS ? 1 : 8], GB[GS ? 1 : 8];
        // GS: the zone holds two streams in halves 0 and 1.  Group order  p complex (x,y | z,r3 of p) -> packed real
        // (z,r3 of p | z,r3 of q) -> q complex (x,y | z,r3 of q): every change of group replaces ONE stream, and that
        // copy is issued as soon as the last series of the old group has formed its samples, so it lands under that
        // series' phase C instead of being waited for (the copies were 11 % of the kernel's stall samples).
        //   half 0: (x,y)_p  -> (z,r3)_q            half 1: (z,r3)_p  -> (x,y)_q   (idx >= 7)
        // roles: idx 0-3  ga = half 0, gb = half 1 | idx 4-6  ga = half 1, gb = half 0 | idx 7-10  ga = half 1, gb = half 0
        // (the roles are compile-time: every kind of series is instantiated for the placement its group uses)
        auto gz = [&](int half, int a) -> cd { return ZG[(8 * half + a) * Shape::NT]; };
        auto copy_stream = [&](const cd* src, int half) {   // 8 copies of this thread's residue into one half of its zone
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const i64 fr = (i64)(B * a + bA) * 4096 + n2A;
                cp_async16_zfill(ZG + (8 * half + a) * Shape::NT, src + fr, (onmask >> a) & 1u);
            }
        };
        if constexpr (GS) {
            if (item == (i64)blockIdx.y) {   // first item of this CTA: nothing was prefetched
                copy_stream(g, 0);
                copy_stream(g + N, 1);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
#pragma unroll 1
        for (int idx = 0; idx < 11; ++idx) {
            int slot;
            if constexpr (GS) {
                // idx 0-3: Y21, Y22, F21, F22 of p | 4-6: Y20, r^-3, F20 (packed) | 7-10: Y21, Y22, F21, F22 of q
                slot = idx < 4 ? (idx == 0 ? 1 : idx == 1 ? 3 : idx == 2 ? 7 : 9)
                     : idx < 7 ? (idx == 4 ? 0 : idx + 0)
                               : (idx == 7 ? 2 : idx == 8 ? 4 : idx == 9 ? 8 : 10);
                if (idx == 0 || idx == 4 || idx == 7) asm volatile("cp.async.wait_group 0;" ::: "memory");
            } else {
                slot = idx < 8 ? ((idx & 3) == 0 ? 1 : (idx & 3) == 1 ? 3 : (idx & 3) == 2 ? 7 : 9) + (idx >> 2)
                               : (idx == 8 ? 0 : idx - 4);
                if (idx == 0 || idx == 4 || idx == 8) {
                    const cd* sa = g + (idx == 0 ? 0 : idx == 4 ? 2 : 1) * N;
                    const cd* sb = g + (idx == 0 ? 1 : 3) * N;
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const i64 fr = (i64)(B * a + bA) * 4096 + n2A;
                        const bool on = (onmask >> a) & 1u;
                        const cd zero = make_double2(0.0, 0.0);
                        GA[a] = on ? ldg_cd(sa + fr) : zero;
                        GB[a] = on ? ldg_cd(sb + fr) : zero;
                    }
                }
            }
            // ---- phase A: the 8 samples of residue b, one specialised block per kind of series (the slot is uniform
            // over the CTA).  Frames beyond the series and a missing second pair arrive as zero geometry, which makes
            // every series with a factor x, y, z or r^-3 vanish; only Y20 and r^-3 - <r^-3> switch their constant off.
            cd v[8];
            auto form = [&](auto kind, auto swapped) {
                constexpr int K = decltype(kind)::value;   // 0 Y20 | 5 r^-3 | 6 F20 | 1 Y21 | 3 Y22 | 7 F21 | 9 F22
                constexpr bool SW = decltype(swapped)::value;   // GS: first stream of the group in half 1, second in half 0
                auto ga = [&](int a) -> cd { if constexpr (GS) return gz(SW ? 1 : 0, a); else return GA[a]; };
                auto gb = [&](int a) -> cd { if constexpr (GS) return gz(SW ? 0 : 1, a); else return GB[a]; };
                if constexpr (K == 0 || K == 5 || K == 6) {   // GA, GB = (z/r, r^-3) of p, q -> re + i im
                    const unsigned onq = q_exists ? onmask : 0u;
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        if constexpr (K == 0) {
                            const cd zp = ga(a), zq = gb(a);
                            v[a].x = (((onmask >> a) & 1u) ? C20 : 0.0) * fma(3.0 * zp.x, zp.x, -1.0);
                            v[a].y = (((onq >> a) & 1u) ? C20 : 0.0) * fma(3.0 * zq.x, zq.x, -1.0);
                        } else if constexpr (K == 5) {
                            v[a].x = ga(a).y - (((onmask >> a) & 1u) ? mp : 0.0);
                            v[a].y = gb(a).y - (((onq >> a) & 1u) ? mq : 0.0);
                        } else {
                            const cd zp = ga(a), zq = gb(a);
                            v[a].x = (KF * C20) * zp.y * fma(3.0 * zp.x, zp.x, -1.0);
                            v[a].y = (KF * C20) * zq.y * fma(3.0 * zq.x, zq.x, -1.0);
                        }
                    }
                } else {                  // GA = (x, y)/r, GB = (z/r, r^-3) of the series' pair
                    constexpr bool weighted = K >= 7;
                    constexpr bool order1 = K == 1 || K == 7;
                    constexpr double c = (order1 ? C21 : C22) * (weighted ? KF : 1.0);
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const cd xy = ga(a), zr = gb(a);
                        const cd w = order1 ? xy : make_double2(xy.x * xy.x - xy.y * xy.y, 2.0 * xy.x * xy.y);
                        const double sc = c * (weighted ? zr.y : 1.0) * (order1 ? zr.x : 1.0);
                        v[a] = make_double2(sc * w.x, sc * w.y);
                    }
                }
            };
            using SwNo = std::integral_constant<bool, false>;
            using SwYes = std::integral_constant<bool, true>;
            if constexpr (GS) {   // p complex: (x,y) in half 0, (z,r3) in half 1; packed and q complex: the other way round
                switch (slot) {
                    case 0: form(TpcKind<0>{}, SwYes{}); break;
                    case 5: form(TpcKind<5>{}, SwYes{}); break;
                    case 6: form(TpcKind<6>{}, SwYes{}); break;
                    case 1: form(TpcKind<1>{}, SwNo{}); break;
                    case 2: form(TpcKind<1>{}, SwYes{}); break;
                    case 3: form(TpcKind<3>{}, SwNo{}); break;
                    case 4: form(TpcKind<3>{}, SwYes{}); break;
                    case 7: form(TpcKind<7>{}, SwNo{}); break;
                    case 8: form(TpcKind<7>{}, SwYes{}); break;
                    case 9: form(TpcKind<9>{}, SwNo{}); break;
                    default: form(TpcKind<9>{}, SwYes{}); break;
                }
            } else {
                switch (slot) {
                    case 0: form(TpcKind<0>{}, SwNo{}); break;
                    case 5: form(TpcKind<5>{}, SwNo{}); break;
                    case 6: form(TpcKind<6>{}, SwNo{}); break;
                    case 1: case 2: form(TpcKind<1>{}, SwNo{}); break;
                    case 3: case 4: form(TpcKind<3>{}, SwNo{}); break;
                    case 7: case 8: form(TpcKind<7>{}, SwNo{}); break;
                    default: form(TpcKind<9>{}, SwNo{}); break;
                }
            }
            if constexpr (!Shape::SPLIT_A) {
                cd e[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) e[a] = v[a];
                fft8(e);        // y[2 c']
                fft8_odd(v);    // y[2 c' + 1]
                smem[bA * C + colA] = e[0];
#pragma unroll
                for (int c = 1; c < 16; ++c) {
                    const cd y = (c & 1) ? v[c >> 1] : e[c >> 1];
                    smem[c * SROW + bA * C + colA] = cmul(y, TA[c * B + bA]);
                }
            } else if (halfA == 0) {
                fft8(v);        // y[2 c']
#pragma unroll
                for (int c = 0; c < 16; c += 2) smem[c * SROW + bA * C + colA] = c ? cmul(v[c >> 1], TA[c * B + bA]) : v[0];
            } else {
                fft8_odd(v);    // y[2 c' + 1]
#pragma unroll
                for (int c = 1; c < 16; c += 2) smem[c * SROW + bA * C + colA] = cmul(v[c >> 1], TA[c * B + bA]);
            }
            __syncthreads();
            if constexpr (GS) {   // phase A is done, no sample registers are live: the stream the next group replaces may be
                                  // overwritten now; the copy lands under phase C
                if (idx == 3) { copy_stream(g + 3 * N, 0); asm volatile("cp.async.commit_group;" ::: "memory"); }
                if (idx == 6) { copy_stream(g + 2 * N, 1); asm volatile("cp.async.commit_group;" ::: "memory"); }
                if (idx == 10 && item + gridDim.y < p.items) {
                    const cd* gn = p.Gs + (item + gridDim.y) * 4 * N;
                    copy_stream(gn, 0);
                    copy_stream(gn + N, 1);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
            }
            // ---- phase C: B-point transform over b for (c, column)
            cd u[HB];
            if constexpr (Shape::SPLIT_C) {   // parity `par`: sums / differences of the two halves, B/2-point codelet
#pragma unroll
                for (int j = 0; j < HB; ++j) {
                    const cd y0 = smem[cC * SROW + j * C + colC];
                    const cd y1 = smem[cC * SROW + (j + HB) * C + colC];
                    u[j] = par ? make_double2(y0.x - y1.x, y0.y - y1.y) : make_double2(y0.x + y1.x, y0.y + y1.y);
                }
            } else {
#pragma unroll
                for (int j = 0; j < HB; ++j) u[j] = smem[cC * SROW + j * C + colC];
            }
            __syncthreads();   // the buffer may be rewritten by the next series
            if constexpr (Shape::SPLIT_C) {
                if (par) Tpc2Codelets<B>::odd(u); else Tpc2Codelets<B>::even(u);
            } else if constexpr (B == 16) {
                fft16(u);
            } else {
                fft8(u);
            }
            cd* dst = p.T + (item * 11 + slot) * TS + n2C;
            {
                cd w = seed;
#pragma unroll
                for (int dd = 0; dd < HB; ++dd) {
                    const int k1 = cC + 16 * par + KSTEP * dd;
                    st_stream(dst + (i64)k1 * DYNPY_T_RS, cmul(u[dd], w));
                    if (dd + 1 < HB) w = cmul(w, step);
                }
            }
        }
    }
}

}  // namespace dynpy
