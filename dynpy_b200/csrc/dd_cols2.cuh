// dynpy_b200 — fused dipolar word generator + first-pass (column) transform, warp-autonomous.
//
// Replaces, for one batch of pair-pairs ("items"), distances.pdist_ortho (minimum image,
// distances.py:190-234), ddrax.cart_to_dipolar_from_df (ddrax.py:110-126) and the first half of
// the zero-padded forward FFT of wiener_khinchin (ddrax.py:276-285) for all 11 real series of a
// pair: nothing but the four-step intermediate T is written.
//
// Decomposition M = M1 x 4096, M1 = 8 R1 (used for M1 = 64; M1 <= 50 goes through the thread-per-column
// kernel of dd_tpc.cuh); sample n = 4096 n1 + n2.
//   * a WARP owns CW = 32/R1 adjacent columns n2 (R1 threads per column, 8 points per thread);
//     the only data exchange of the M1-point column transform is inside the warp (4.5 KB of
//     shared memory per warp, __syncwarp), there is no CTA-wide barrier in the kernel;
//   * a thread computes the pair geometry (minimum image per axis, unit vector, r^-3) of its own
//     4 frames once per pair and derives all series of the pair from it in registers;
//   * step 1: radix-R1 codelet on the zero-padded half (n1 >= M1/2 is structurally zero),
//     twiddle W_M1^{m k}; step 2: radix-8 codelet, inter-pass twiddle W_M^{n2 k1} from 8 complex
//     registers per thread (fixed while the warp stays on its columns), 16-byte stores to
//     T[item][slot][k1][n2];
//   * the work list (item, 16-column tile) is cut into equal contiguous ranges per CTA, item-major, so
//     the whole grid writes into the T blocks of the same few items at any time (TLB reach); the
//     inter-pass twiddles of a tile are 8 table look-ups per thread (table filled once per call).
// The mean of r^-3 (G_r is the ACF of r^-3 - <r^-3>, ddrax.py:292) is NOT subtracted here: the
// per-pair sums go to rsum[] and the row kernel subtracts mean x (column transform of the
// all-ones series), which is linear and exact; `ones` mode produces that table.
#pragma once
#include "codelets.cuh"
#include "dd_device.cuh"
#include "fft_engine.cuh"
#include "fft_rows2.cuh"  // DYNPY_T_RS: layout of T

namespace dynpy {

struct ColsDD2Params {
    const double* X;   // wrapped coordinates [atom][3][N]
    const int* pi;
    const int* pj;
    i64 pair0;         // first pair of the batch (even)
    i64 n_pairs;       // total number of pairs
    i64 items;         // pair-pairs in the batch
    i64 N;
    double a, b, c;
    cd* T;             // [items][11][M1][DYNPY_T_RS]   (ones mode: [M1][DYNPY_T_RS])
    i64 M;
    double inv_M;
    double* rsum;      // [2 * items] sums of r^-3 (zeroed by the caller)
    const cd* twp;     // W_4096^j
    const cd* twt;     // inter-pass twiddles W_M^{n2 k1}, [M1][4096]
    int ones;
};

#define DD2_MAX_IDX 256  // pairs whose atom indices a CTA caches in shared memory

template <int R> struct HalfCodelet;
template <> struct HalfCodelet<2> { static __device__ __forceinline__ void run(cd (&v)[2]) { fft2_h1(v); } };
template <> struct HalfCodelet<4> { static __device__ __forceinline__ void run(cd (&v)[4]) { fft4_h2(v); } };
template <> struct HalfCodelet<8> { static __device__ __forceinline__ void run(cd (&v)[8]) { fft8_h4(v); } };

__device__ __forceinline__ void cp_async8_(double* smem_dst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");  // 8-byte copies exist only as .ca
}

// compile-time shape of the column kernel for a given M1 = 8 R1
template <int M1> struct ColShape {
    static constexpr int TPC = M1 / 8;              // threads per column == first radix R1
    static constexpr int CW = 32 / TPC;             // columns per warp
    static constexpr int NFR = 4;                   // frames (non-zero samples) per thread
    static constexpr int NTW = 8;                   // outputs == inter-pass twiddles per thread
    static constexpr int NWARP = 4;                 // warps per CTA; the CTA owns NWARP*CW adjacent columns
    static constexpr int TCOLS = CW * NWARP;
    static constexpr int NTILE = 4096 / TCOLS;
    static constexpr int WSLOTS = (M1 + M1 / 8) * CW;  // exchange buffer of one warp (cd units)
    static constexpr int NTW1 = 64;                 // step-1 twiddle table entries
    static constexpr size_t smem_bytes() {
        return (size_t)NWARP * WSLOTS * sizeof(cd) + (size_t)NFR * 6 * 128 * sizeof(double) + (size_t)NTW * 128 * sizeof(cd) +
               2 * DD2_MAX_IDX * sizeof(int) + NTW1 * sizeof(cd);
    }
};

template <int M1>
__global__ void __launch_bounds__(128, 3) k_cols_dd2(const ColsDD2Params p) {
    typedef ColShape<M1> SH;
    constexpr int R1 = SH::TPC, CW = SH::CW, NFR = SH::NFR, NTW = SH::NTW, NWARP = SH::NWARP, NTILE = SH::NTILE;
    constexpr int WSLOTS = SH::WSLOTS;
    constexpr int QH = R1 / 2;  // non-zero inputs of a step-1 butterfly
    constexpr int NIT = 8 / R1; // step-1 butterflies per thread
    extern __shared__ double2 smem[];
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int c = lane % CW;
    const int bb = lane / CW;
    const bool lane_on = bb < R1;
    cd* sm = smem + wid * WSLOTS;
    // thread-private staging of the NEXT pair's coordinates: value v of this thread at stage[v*128] ...
    double* stage = reinterpret_cast<double*>(smem + NWARP * WSLOTS) + threadIdx.x;
    // ... of the next tile's inter-pass twiddles: stw[j*128] ...
    cd* stw = smem + NWARP * WSLOTS + (NFR * 6 * 128) / 2 + threadIdx.x;
    // ... the atom indices of the pairs this CTA will visit (filled once) ...
    int* sidx = reinterpret_cast<int*>(smem + NWARP * WSLOTS + (NFR * 6 * 128) / 2 + NTW * 128);
    // ... and the step-1 twiddles W_M1^{m k} as [k][m]
    cd* stw1 = smem + NWARP * WSLOTS + (NFR * 6 * 128) / 2 + NTW * 128 + (2 * DD2_MAX_IDX * (int)sizeof(int)) / (int)sizeof(cd);
    if (threadIdx.x < 64) stw1[threadIdx.x] = ldg_cd(p.twp + (4096 / M1) * (threadIdx.x & 7) * (threadIdx.x >> 3));
    auto slot_of = [&](int row) { return (row + (row >> 3)) * CW + c; };

    const int items = p.ones ? 1 : (int)p.items;
    const int units = NTILE * items;
    const int w0 = (int)((i64)units * blockIdx.x / gridDim.x), w1 = (int)((i64)units * (blockIdx.x + 1) / gridDim.x);
    if (w0 >= w1) return;
    const int nfr = (int)(p.N < (i64)M1 * 2048 ? p.N : (i64)M1 * 2048);  // frames beyond M/2 never exist
    constexpr i64 TS = (i64)M1 * DYNPY_T_RS;  // one transform of T

    const double C20 = 0.31539156525252000603, C21 = -0.77254840404638362, C22 = 0.38627420202319181;
    const double KF = 1.5853309190424043;

    int n2 = 0;
    bool on = false;  // this thread owns a real column (lane in use and n2 < 4096)
    cd tw[NTW];
    // n1 of this thread's fs-th frame, and k1 of its j-th output
    auto n1_of = [&](int fs) -> int { return (fs % QH) * 8 + bb + (fs / QH) * R1; };
    auto k1_of = [&](int j) -> int { return bb + R1 * j; };
    auto col_of = [&](int tile) -> int { return (tile * NWARP + wid) * CW + c; };
    // one series of the current (item, slot): NFR samples per thread -> column transform -> T
    auto emit = [&](const cd (&smp)[NFR], cd* dst) {
        {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int m = bb + it * R1;
                cd u[R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) u[q] = (q < QH) ? smp[it * QH + q] : make_double2(0.0, 0.0);
                HalfCodelet<R1>::run(u);
                sm[slot_of(m)] = u[0];
#pragma unroll
                for (int k = 1; k < R1; ++k) sm[slot_of(k * 8 + m)] = cmul(u[k], stw1[k * 8 + m]);
            }
            __syncwarp();
            cd u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) u[q] = sm[slot_of(bb * 8 + q)];
            __syncwarp();
            fft8(u);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[(bb + R1 * j) * DYNPY_T_RS] = cmul(u[j], tw[j]);
        }
    };
    auto set_tile = [&](int tile, bool staged) {
        n2 = col_of(tile);
        on = lane_on && n2 < 4096;
#pragma unroll
        for (int j = 0; j < NTW; ++j)
            tw[j] = staged ? stw[j * 128] : (on ? ldg_cd(p.twt + ((i64)k1_of(j) << 12) + n2) : make_double2(0.0, 0.0));
    };
    auto t_off = [&]() -> i64 { return n2; };
    __syncthreads();  // stw1 is filled

    if (p.ones) {
        for (int w = w0; w < w1; ++w) {
            set_tile(w % NTILE, false);
            cd smp[NFR];
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) smp[fs] = make_double2(on && (n1_of(fs) << 12) + n2 < nfr ? 1.0 : 0.0, 0.0);
            emit(smp, p.T + t_off());
        }
        return;
    }

    // the warp walks "pair steps" s = 2 (unit - w0) + hq; coordinates of step s+1 are staged while
    // step s is transformed, its atom indices are fetched one step earlier still
    const int S = 2 * (w1 - w0);
    // atom indices of the items [item_lo, item_hi] this CTA touches -> shared memory (the index arrays
    // are evicted from L2 by the T stream; a dependent global look-up per pair costs microseconds)
    const int item_lo = w0 / NTILE, item_hi = (w1 - 1) / NTILE;
    const int n_idx = 2 * (item_hi - item_lo + 1);
    for (int k = threadIdx.x; k < n_idx && k < DD2_MAX_IDX; k += blockDim.x) {
        const i64 pr = p.pair0 + 2 * (i64)item_lo + k;
        const bool ok = pr < p.n_pairs;
        sidx[2 * k] = ok ? __ldg(p.pi + pr) : -1;
        sidx[2 * k + 1] = ok ? __ldg(p.pj + pr) : -1;
    }
    __syncthreads();
    auto atoms_of = [&](int s, int& ai, int& aj) {
        ai = -1; aj = -1;
        if (s < S) {
            const int k = 2 * ((w0 + (s >> 1)) / NTILE - item_lo) + (s & 1);
            if (k < DD2_MAX_IDX) { ai = sidx[2 * k]; aj = sidx[2 * k + 1]; }
            else {
                const i64 pr = p.pair0 + 2 * (i64)item_lo + k;
                if (pr < p.n_pairs) { ai = __ldg(p.pi + pr); aj = __ldg(p.pj + pr); }
            }
        }
    };
    auto stage_step = [&](int s, int ai, int aj) {
        if (s < S) {
            const int col = col_of((w0 + (s >> 1)) % NTILE);
            const bool col_on = lane_on && col < 4096;
            if (ai >= 0 && col_on) {
                const double* xi = p.X + (i64)ai * 3 * p.N;
                const double* xj = p.X + (i64)aj * 3 * p.N;
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) {
                    const int fr = (n1_of(fs) << 12) + col;
                    if (fr < nfr) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            cp_async8_(stage + (fs * 6 + d) * 128, xi + d * p.N + fr);
                            cp_async8_(stage + (fs * 6 + 3 + d) * 128, xj + d * p.N + fr);
                        }
                    }
                }
            }
            if ((s & 1) == 0 && col_on) {  // first pair of a unit: its tile's inter-pass twiddles
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(stw + j * 128);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(p.twt + ((i64)k1_of(j) << 12) + col) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int ai, aj, ai_n, aj_n;
    atoms_of(0, ai, aj);
    stage_step(0, ai, aj);
    atoms_of(1, ai_n, aj_n);

    double y20p[NFR], r3p[NFR];
    for (int s = 0; s < S; ++s) {
        const int unit = w0 + (s >> 1);
        const int item = unit / NTILE;  // item-major: the whole grid works on the same few items at any time,
        const int tile = unit - item * NTILE;  // which keeps the pages of T and of the coordinates in the TLBs
        const int hq = s & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (hq == 0) set_tile(tile, true);
        cd* tbase = p.T + (i64)item * 11 * TS + t_off();
        const bool exists = ai >= 0 && on;
        double gx[NFR], gy[NFR], gz[NFR], g3[NFR];
        {
            double rs = 0.0;
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) {
                const bool ok = exists && (n1_of(fs) << 12) + n2 < nfr;
                double v[6];
#pragma unroll
                for (int d = 0; d < 6; ++d) v[d] = ok ? stage[(fs * 6 + d) * 128] : (d < 3 ? 0.0 : 1.0);
                double dx, dy, dz, sx, sy, sz;
                axis_min(v[0], v[3], p.a, dx, sx);
                axis_min(v[1], v[4], p.b, dy, sy);
                axis_min(v[2], v[5], p.c, dz, sz);
                const double r2 = __dadd_rn(__dadd_rn(sx, sy), sz);
                double rinv = rsqrt(r2);
                rinv = rinv * fma(-0.5 * r2, rinv * rinv, 1.5);
                rinv = ok ? rinv : 0.0;
                gx[fs] = dx * rinv; gy[fs] = dy * rinv; gz[fs] = dz * rinv;
                g3[fs] = rinv * rinv * rinv;
                rs += g3[fs];
            }
            // the staging slots are free again: next step's coordinates, and the atoms of the one after
            const bool pair_exists = ai >= 0;
            ai = ai_n; aj = aj_n;
            stage_step(s + 1, ai, aj);
            atoms_of(s + 2, ai_n, aj_n);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rs += __shfl_down_sync(0xffffffffu, rs, o);
            if (lane == 0 && pair_exists) atomicAdd(p.rsum + 2 * item + hq, rs);
        }
        cd smp[NFR];
#pragma unroll
        for (int fs = 0; fs < NFR; ++fs) {
            const double k = C21 * gz[fs];
            smp[fs] = make_double2(k * gx[fs], k * gy[fs]);
        }
        emit(smp, tbase + (1 + hq) * TS);
#pragma unroll
        for (int fs = 0; fs < NFR; ++fs) {
            const double k = (KF * C21) * g3[fs] * gz[fs];
            smp[fs] = make_double2(k * gx[fs], k * gy[fs]);
        }
        emit(smp, tbase + (7 + hq) * TS);
#pragma unroll
        for (int fs = 0; fs < NFR; ++fs)
            smp[fs] = make_double2(C22 * (gx[fs] * gx[fs] - gy[fs] * gy[fs]), (2.0 * C22) * gx[fs] * gy[fs]);
        emit(smp, tbase + (3 + hq) * TS);
#pragma unroll
        for (int fs = 0; fs < NFR; ++fs) {
            const double k = KF * g3[fs];
            smp[fs] = make_double2(k * smp[fs].x, k * smp[fs].y);
        }
        emit(smp, tbase + (9 + hq) * TS);
        if (hq == 0) {
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) {
                const bool ok = exists && (n1_of(fs) << 12) + n2 < nfr;
                y20p[fs] = ok ? C20 * fma(3.0 * gz[fs], gz[fs], -1.0) : 0.0;
                r3p[fs] = g3[fs];
            }
        } else {
            double y20q[NFR];
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) {
                const bool ok = exists && (n1_of(fs) << 12) + n2 < nfr;
                y20q[fs] = ok ? C20 * fma(3.0 * gz[fs], gz[fs], -1.0) : 0.0;
                smp[fs] = make_double2(y20p[fs], y20q[fs]);
            }
            emit(smp, tbase);
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) smp[fs] = make_double2(KF * r3p[fs] * y20p[fs], KF * g3[fs] * y20q[fs]);
            emit(smp, tbase + 6 * TS);
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) smp[fs] = make_double2(r3p[fs], g3[fs]);
            emit(smp, tbase + 5 * TS);
        }
    }
}

// inter-pass twiddle table  twt[k1][n2] = W_M^{n2 k1}
__global__ void k_fill_twt(cd* twt, int M1, i64 M, double inv_M) {
    const i64 n = (i64)M1 << 12;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        twt[i] = w_exact(((i >> 12) * (i & 4095)) % M, inv_M);
}

}  // namespace dynpy
