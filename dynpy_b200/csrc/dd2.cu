// Host side of the dipolar fast path: fused generator + column kernel (dd_cols2.cuh) and the
// accumulate-power row kernel (fft_rows2.cuh), batched through the T workspace.
#include "dd_cols2.cuh"
#include "dd_tpc.cuh"
#include "fft_rows2.cuh"
#include "kernels.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace dynpy {

static inline i64 align_up2(i64 x, i64 a) { return (x + a - 1) / a * a; }

bool dd2_supported(i64 M) {
    const i64 m1 = M >> 12;
    return (M & 4095) == 0 && (m1 == 16 || m1 == 20 || m1 == 24 || m1 == 32 || m1 == 40 || m1 == 50 || m1 == 64 ||
                               m1 == 128 || m1 == 256 || m1 == 512);
}

// bytes of one padded transform of T
static inline i64 t_bytes(i64 M) { return (M >> 12) * (i64)DYNPY_T_RS * (i64)sizeof(cd); }

// workspace: R[M'] | TWT[M] | rsum[2*cap] | Gs[2*cap][4][N] | T[cap][11][M']   (M' = M1 * DYNPY_T_RS padded rows)
i64 dd2_workspace_bytes(i64 M, i64 N, i64 items) {
    if (items < 1) items = 1;
    return t_bytes(M) + M * (i64)sizeof(cd) + align_up2(items * 16, 256) + align_up2(items * 64 * N, 256) +
           items * DD_NSLOTS * t_bytes(M) + 1024;
}

// long columns (M1 = 128, 256, 512): two-level thread-per-column kernel, C columns per CTA
#ifndef TPC2_C
#define TPC2_C 4
#endif
template <int B, int C, int PER, bool GS>
static cudaError_t launch_cols_tpc3_shape(const ColsTpcParams& p, const cd* tw, int n_sm, cudaStream_t st) {
    const size_t smem3 = Tpc3Shape<B, C>::smem_bytes(GS);
    auto kern = k_cols_tpc3<B, C, PER, GS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
    if (e != cudaSuccess) return e;
    const int tiles = 4096 / C;
    int gy = (4 * PER * n_sm + tiles - 1) / tiles;   // >= 4 waves
    if (gy > p.items) gy = (int)p.items;
    if (gy < 1) gy = 1;
    kern<<<dim3(tiles, gy), Tpc3Shape<B, C>::NT, smem3, st>>>(p, tw);
    return cudaGetLastError();
}
template <int B, int C>
static cudaError_t launch_cols_tpc3_B(const ColsTpcParams& p, const cd* tw, int n_sm, cudaStream_t st, int per, int gs) {
    // per = CTAs per SM = register budget (2: 255, 3: 168, 4: 128); gs: geometry staged in shared memory
    if (gs) return per == 2 ? launch_cols_tpc3_shape<B, C, 2, true>(p, tw, n_sm, st)
                 : per == 4 ? launch_cols_tpc3_shape<B, C, 4, true>(p, tw, n_sm, st)
                            : launch_cols_tpc3_shape<B, C, 3, true>(p, tw, n_sm, st);
    return per == 3 ? launch_cols_tpc3_shape<B, C, 3, false>(p, tw, n_sm, st)
         : per == 4 ? launch_cols_tpc3_shape<B, C, 4, false>(p, tw, n_sm, st)
                    : launch_cols_tpc3_shape<B, C, 2, false>(p, tw, n_sm, st);
}
static cudaError_t launch_cols_tpc2(const ColsTpcParams& p, const cd* tw, int n_sm, cudaStream_t st) {
    ProfScope prof(PROF_COLS, st);
    const int m1 = (int)(p.M >> 12);
    const char* pe = getenv("DYNPY_TPC3_PER_SM");   // A/B knobs: CTAs per SM, geometry staging (DYNPY_TPC3_GS=0|1)
    const char* ge = getenv("DYNPY_TPC3_GS");
    // measured (96 pairs, tools/dbg/probe_dd.py): 512-point columns 6.08 ms with the geometry in registers at 2 CTAs
    // per SM, 5.60 staged in shared memory at 3; 256-point: 2.39 / 2.28; 128-point: registers at 3 CTAs (1.13 vs 1.23)
    const int gs = ge ? atoi(ge) : (m1 == 128 ? 0 : 1);
    const int per = pe ? atoi(pe) : 3;
    if (m1 == 256) return launch_cols_tpc3_B<16, 8>(p, tw, n_sm, st, per, gs);   // 128 threads, fft16 per phase-C thread
    if (m1 == 128) return launch_cols_tpc3_B<8, 8>(p, tw, n_sm, st, per, gs);    // 128 threads, fft8 per phase-C thread
    constexpr int B = 32, C = TPC2_C;
    const size_t smem = (size_t)16 * B * C * sizeof(cd);
    const char* en = getenv("DYNPY_TPC2");
    const int v3 = (en && en[0] == 'o') ? 0 : 1;   // "old": per-series loads (A/B)
    if (v3) return launch_cols_tpc3_B<B, C>(p, tw, n_sm, st, per, gs);
    cudaError_t e = cudaFuncSetAttribute(k_cols_tpc2<B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tiles = 4096 / C;
    const int per_sm = 384 / (B * C);
    int gy = (4 * per_sm * n_sm + tiles - 1) / tiles;   // >= 4 waves
    if (gy > p.items) gy = (int)p.items;
    if (gy < 1) gy = 1;
    k_cols_tpc2<B, C><<<dim3(tiles, gy), B * C, smem, st>>>(p, tw);
    return cudaGetLastError();
}

static inline dim3 geom_grid(i64 chunks, i64 slots, int pair_fastest) {
    return pair_fastest ? dim3((unsigned)slots, (unsigned)chunks) : dim3((unsigned)chunks, (unsigned)slots);
}

template <int M1>
static cudaError_t launch_cols_dd2_M1(const ColsDD2Params& p, int n_sm, cudaStream_t st) {
    auto kern = k_cols_dd2<M1>;
    const size_t smem = ColShape<M1>::smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ProfScope prof(PROF_COLS, st);
    kern<<<3 * n_sm, 128, smem, st>>>(p);
    return cudaGetLastError();
}

static cudaError_t launch_cols_dd2(const ColsDD2Params& p, int n_sm, cudaStream_t st) {
    switch ((int)(p.M >> 12)) {
        case 32: return launch_cols_dd2_M1<32>(p, n_sm, st);  // A/B checks only (DYNPY_DD_COLS=warp)
        case 64: return launch_cols_dd2_M1<64>(p, n_sm, st);
        default: return cudaErrorInvalidValue;
    }
}

// rows of the slots listed in sel[0..nsel); `fix` selects the mean-removing instantiation
cudaError_t launch_rows2(Rows2Params p, const int* sel, int nsel, bool fix, int n_sm, cudaStream_t st) {
    const size_t smem = (size_t)(pad16(4096) + 256) * sizeof(cd);  // row buffer + step-1 twiddle table
    p.nsel = nsel;
    p.sel_pack = 0;
    for (int s = 0; s < nsel; ++s) p.sel_pack |= (unsigned long long)sel[s] << (4 * s);
    const i64 units = (i64)nsel * p.M1 * p.items;
    if (units <= 0) return cudaSuccess;
    {
        static int pf = -1;
        if (pf < 0) { const char* e = getenv("DYNPY_ROWS_PF"); pf = e ? atoi(e) : 1; }
        p.pf_dist = pf;
    }
    int grid = 2 * n_sm;
    if (grid > units) grid = (int)units;
    ProfScope prof(PROF_ROWS, st);
    cudaError_t e;
    if (fix) {
        e = cudaFuncSetAttribute(k_rows2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_rows2<true><<<grid, 256, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_rows2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_rows2<false><<<grid, 256, smem, st>>>(p);
    }
    return cudaGetLastError();
}

// 4096-point rows with a write epilogue (fft_engine.cuh EpiMode 1..3) through the same row kernel: T holds
// transforms f_base .. f_base + n - 1, one after the other
cudaError_t launch_rows2_write(Rows2Params p, int mode, int n_sm, cudaStream_t st) {
    const size_t smem_row = (size_t)(pad16(4096) + 256) * sizeof(cd);
    const size_t smem = mode == ROWS2_WRITE_POWER ? smem_row + (size_t)pad16(4096) * sizeof(double) : smem_row;
    const i64 units = (i64)p.M1 * p.items;
    if (units <= 0) return cudaSuccess;
    p.nsel = 1;
    p.sel_pack = 0;
    p.nslots = 1;
    {
        static int pf = -1;
        if (pf < 0) { const char* e = getenv("DYNPY_ROWS_PF"); pf = e ? atoi(e) : 1; }
        p.pf_dist = pf;
    }
    int grid = 2 * n_sm;
    if (grid > units) grid = (int)units;
    ProfScope prof(PROF_ROWS, st);
    cudaError_t e;
#define ROWS2_WRITE_CASE(MODE)                                                                                          \
    e = cudaFuncSetAttribute(k_rows2<false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e != cudaSuccess) return e;                                                                                    \
    k_rows2<false, MODE><<<grid, 256, smem, st>>>(p)
    switch (mode) {
        case ROWS2_WRITE_REAL: ROWS2_WRITE_CASE(ROWS2_WRITE_REAL); break;
        case ROWS2_WRITE_POWER: ROWS2_WRITE_CASE(ROWS2_WRITE_POWER); break;
        case ROWS2_WRITE_REAL2: ROWS2_WRITE_CASE(ROWS2_WRITE_REAL2); break;
        case ROWS2_WRITE_INPLACE: ROWS2_WRITE_CASE(ROWS2_WRITE_INPLACE); break;
        default: return cudaErrorInvalidValue;
    }
#undef ROWS2_WRITE_CASE
    return cudaGetLastError();
}

// T[f][k1][k2] (complex rows written back by k_rows2<.., ROWS2_WRITE_INPLACE>) -> real outputs in natural order
// k = k1 + M1 k2 < n_out:  planes = 1: out[f_base + f][k] = scale Re;  planes = 2: out[2 (f_base + f)][k] = scale Re,
// out[2 (f_base + f) + 1][k] = scale Im (while that row exists).  Tiled transpose: reads are 512-byte runs along k2,
// writes 256-byte runs along k1.
__global__ void __launch_bounds__(256) k_rows_to_natural(const cd* __restrict__ T, double* __restrict__ out, int M1, int rs,
                                                         i64 f_base, i64 n_out, i64 n_series, int planes, double scale) {
    __shared__ cd tile[32][33];
    const i64 f = blockIdx.z;
    const cd* in = T + f * (i64)M1 * rs;
    const int k2_0 = blockIdx.x * 32, k1_0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int k1 = k1_0 + j;
        if (k1 < M1) tile[j][tx] = in[(i64)k1 * rs + k2_0 + tx];
    }
    __syncthreads();
    const i64 fo = f_base + f;
    double* o0 = out + (planes == 2 ? 2 * fo : fo) * n_out;
    const bool two = planes == 2 && 2 * fo + 1 < n_series;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int k1 = k1_0 + tx;
        const i64 k = (i64)k1 + (i64)M1 * (k2_0 + j);
        if (k1 < M1 && k < n_out) {
            const cd v = tile[tx][j];
            o0[k] = v.x * scale;
            if (two) o0[n_out + k] = v.y * scale;
        }
    }
}
cudaError_t launch_rows_to_natural(const cd* T, double* out, int M1, int rs, i64 n_fft, i64 f_base, i64 n_out, i64 n_series,
                                   int planes, double scale, cudaStream_t st) {
    if (n_fft <= 0 || n_out <= 0) return cudaSuccess;
    const i64 k2_end = (n_out + M1 - 1) / M1;   // k2 values that hold any k < n_out
    const unsigned gx = (unsigned)((k2_end + 31) / 32 < 128 ? (k2_end + 31) / 32 : 128);
    ProfScope prof(PROF_ROWS, st);
    for (i64 f0 = 0; f0 < n_fft; f0 += 32768) {
        const i64 cnt = n_fft - f0 < 32768 ? n_fft - f0 : 32768;
        dim3 grid(gx, (unsigned)((M1 + 31) / 32), (unsigned)cnt);
        k_rows_to_natural<<<grid, 256, 0, st>>>(T + f0 * (i64)M1 * rs, out, M1, rs, f_base + f0, n_out, n_series, planes,
                                                scale);
    }
    return cudaGetLastError();
}

template <int M1>
static cudaError_t launch_cols_tpc_M1(const ColsTpcParams& p, int n_sm, cudaStream_t st) {
    ProfScope prof(PROF_COLS, st);
    constexpr int TPC_THREADS = TpcShape<M1>::NT;
    const size_t smem = (size_t)((TPC_DUAL && M1 < 64) ? M1 : M1 / 2) * TPC_THREADS * sizeof(cd);  // per-thread landing zone / sample stash
    cudaError_t e = cudaFuncSetAttribute(k_cols_tpc<M1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_sm = TpcShape<M1>::PER_SM;
    k_cols_tpc<M1><<<per_sm * n_sm, TPC_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}
// noinline: inlined into dd2_run, gcc 13 (-O3) unswitched the batch loop on the alignment test of the geometry branch
// and dropped the cases 16..64 of this switch from the unaligned copy — every N % 4 != 0 call then returned
// cudaErrorInvalidValue from the default branch (found by test_dd_geometry_kernel_variants_agree[99998])
static __attribute__((noinline)) cudaError_t launch_cols_tpc(const ColsTpcParams& p, const cd* tw, int n_sm, cudaStream_t st) {
    switch ((int)(p.M >> 12)) {
        case 16: return launch_cols_tpc_M1<16>(p, n_sm, st);
        case 32: return launch_cols_tpc_M1<32>(p, n_sm, st);
        case 64: return launch_cols_tpc_M1<64>(p, n_sm, st);
        case 50: return launch_cols_tpc_M1<50>(p, n_sm, st);
        case 20: return launch_cols_tpc_M1<20>(p, n_sm, st);
        case 24: return launch_cols_tpc_M1<24>(p, n_sm, st);
        case 40: return launch_cols_tpc_M1<40>(p, n_sm, st);
        case 128: case 256: case 512: return launch_cols_tpc2(p, tw, n_sm, st);
        default: return cudaErrorInvalidValue;
    }
}

int dd2_run(const double* X, i64 N, const int* pi, const int* pj, i64 P, double a, double b, double c, double* S,
            i64 M, void* ws, i64 ws_bytes, const cd* tw, int n_sm, cudaStream_t st) {
    // column kernel: thread-per-column (dd_tpc.cuh); DYNPY_DD_COLS=warp selects the warp-exchange kernel of
    // dd_cols2.cuh where it exists (M1 = 32, 64; A/B checks — it was the faster one for M1 = 64 until the
    // thread-per-column kernel got 255 registers there)
    const char* sel = getenv("DYNPY_DD_COLS");
    const i64 m1 = M >> 12;
    const bool warp_ok = m1 == 32 || m1 == 64;
    const bool tpc = sel ? (sel[0] != 'w' || !warp_ok) : true;
    const i64 items_all = (P + 1) / 2;
    const i64 fixed = t_bytes(M) + M * (i64)sizeof(cd) + 1024;
    const i64 per_item = (i64)DD_NSLOTS * t_bytes(M) + 16 + 64 * N;
    // the largest batch whose carving (same formula as dd2_workspace_bytes) fits the caller's workspace
    i64 cap = (ws_bytes - fixed) / per_item;
    if (cap > items_all) cap = items_all;
    while (cap >= 1 && dd2_workspace_bytes(M, N, cap) > ws_bytes) --cap;
    if (cap < 1) {
        set_error("dipolar fast path: workspace too small (need >= %lld bytes)", (long long)dd2_workspace_bytes(M, N, 1));
        return DYNPY_ERR_WORKSPACE;
    }
    if (cap > items_all) cap = items_all;
    char* base = (char*)ws;
    cd* R = (cd*)base;
    cd* TWT = (cd*)(base + t_bytes(M));
    i64 off = t_bytes(M) + M * (i64)sizeof(cd);
    double* rsum = (double*)(base + off);
    off += align_up2(cap * 16, 256);
    cd* Gs = (cd*)(base + off);
    off += align_up2(cap * 64 * N, 256);
    cd* T = (cd*)(base + off);

    ColsDD2Params cp;
    memset(&cp, 0, sizeof(cp));
    cp.X = X; cp.pi = pi; cp.pj = pj; cp.n_pairs = P; cp.N = N; cp.a = a; cp.b = b; cp.c = c;
    cp.M = M; cp.inv_M = 1.0 / (double)M; cp.twp = tw; cp.rsum = rsum;
    ColsTpcParams tp;
    memset(&tp, 0, sizeof(tp));
    { const char* e = getenv("DYNPY_TPC_SYNC"); tp.sync_units = e ? atoi(e) : 1; }
    i64 tpc_sub = 96;
    { const char* e = getenv("DYNPY_TPC_SUB"); if (e) tpc_sub = atoll(e); }
    tp.N = N; tp.M = M; tp.rsum = rsum; tp.inv_n = 1.0 / (double)N; tp.seeds = TWT;  // the seed table shares the twiddle-table region (never both in one call)
    GeomParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.X = X; gp.pi = pi; gp.pj = pj; gp.n_pairs = P; gp.N = N; gp.a = a; gp.b = b; gp.c = c; gp.Gs = Gs; gp.rsum = rsum;
    { const char* e = getenv("DYNPY_GEOM_ORDER"); gp.pair_fastest = (e && e[0] == 'f') ? 0 : 1; }   // "frames": A/B
    // column transform of the all-ones series (for the mean removal of G_r)
    if (tpc) {
        // the thread-per-column kernel removes the mean of r^-3 in the time domain (the sums are complete
        // when it starts), so there is no all-ones transform and no mean-removing row launch
        k_fill_seeds<<<16, 256, 0, st>>>(TWT, 1.0 / (double)M);
        DYNPY_CHECK(cudaGetLastError(), "dd twiddle seeds");
        tp.Gs = Gs; tp.T = T;
    } else {
        k_fill_twt<<<4 * n_sm, 256, 0, st>>>(TWT, (int)(M >> 12), M, 1.0 / (double)M);
        DYNPY_CHECK(cudaGetLastError(), "dd twiddle table");
        cp.twt = TWT;
        cp.T = R; cp.items = 1; cp.ones = 1;
        DYNPY_CHECK(launch_cols_dd2(cp, n_sm, st), "dd ones columns");
        cp.T = T; cp.ones = 0;
    }

    Rows2Params rp;
    memset(&rp, 0, sizeof(rp));
    rp.T = T; rp.S = S; rp.M = M; rp.M1 = (int)(M >> 12); rp.rs = DYNPY_T_RS; rp.nslots = DD_NSLOTS; rp.twp = tw;
    const int acc_slot[DD_NSLOTS] = {0, 1, 1, 2, 2, 3, 4, 5, 5, 6, 6};
    for (int s = 0; s < DD_NSLOTS; ++s) rp.acc_pack |= (unsigned long long)acc_slot[s] << (4 * s);
    rp.R = R; rp.rsum = rsum; rp.fix_slot = 5; rp.inv_n = 1.0 / (double)N;

    for (i64 it0 = 0; it0 < items_all; it0 += cap) {
        const i64 cnt = items_all - it0 < cap ? items_all - it0 : cap;
        DYNPY_CHECK(cudaMemsetAsync(rsum, 0, sizeof(double) * 2 * cnt, st), "dd rsum");
        if (tpc) {
            gp.pair0 = 2 * it0;
            gp.n_slots = 2 * cnt;
            {
                ProfScope prof(PROF_WORDS, st);
                // N % 4 == 0: 32-byte loads / stores (every stream is 32-byte aligned then); DYNPY_DD_GEOM=scalar for A/B
                const char* ge = getenv("DYNPY_DD_GEOM");
                const int vec = (ge && ge[0] == 's') ? 0 : (ge && ge[0] == 't') ? 2 : 1;
                if (vec == 2 && (N & 1) == 0 && ((uintptr_t)X & 15) == 0) {
                    // TMA-staged form (DYNPY_DD_GEOM=tma): bulk copies through a 4-stage shared-memory ring
                    const size_t smem = (size_t)GEOM_STAGES * 6 * GEOM_TILE * 8 + GEOM_STAGES * 8;
                    DYNPY_CHECK(cudaFuncSetAttribute(k_dd_geom_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "dd geometry (tma) attribute");
                    k_dd_geom_tma<<<2 * n_sm, GEOM_NT, smem, st>>>(gp);
                } else if (vec && (N & 3) == 0 && ((uintptr_t)X & 31) == 0)
                    k_dd_geom4<<<geom_grid((N + 2047) / 2048, 2 * cnt, gp.pair_fastest), 256, 0, st>>>(gp);
                else
                    k_dd_geom<<<geom_grid((N + 1023) / 1024, 2 * cnt, gp.pair_fastest), 256, 0, st>>>(gp);
            }
            DYNPY_CHECK(cudaGetLastError(), "dd geometry");
            // The column kernel deals its units round-robin; over the ~500 rounds of a 32 GiB batch its CTAs drift apart
            // and the 11 series of an (item, column block) meet their geometry in L2 less often (ncu, 820 items per
            // launch: 17.7 MB of geometry re-read from DRAM per item against 7.6 MB at 93 items, ideal 6.4).  The
            // columns of a batch therefore go out in launches of `sub` items (DYNPY_TPC_SUB, 0 = the whole batch) while
            // geometry and rows keep the whole batch: column pass 37.15 -> 36.5 ms per 8192 pairs (sub = 96 or 192).
            tp.n_pairs = P;
            const bool single = (M >> 12) > 64;   // the long-column kernels walk tiles, not units: one launch
            const i64 sub = (single || tpc_sub <= 0) ? cnt : tpc_sub;
            for (i64 s0 = 0; s0 < cnt; s0 += sub) {
                tp.items = cnt - s0 < sub ? cnt - s0 : sub;
                tp.pair0 = 2 * (it0 + s0);
                tp.Gs = Gs + s0 * 4 * N;
                tp.T = T + s0 * (i64)DD_NSLOTS * (t_bytes(M) / (i64)sizeof(cd));
                tp.rsum = rsum + 2 * s0;
                DYNPY_CHECK(launch_cols_tpc(tp, tw, n_sm, st), "dd columns");
            }
        } else {
            cp.pair0 = 2 * it0;
            cp.items = cnt;
            DYNPY_CHECK(launch_cols_dd2(cp, n_sm, st), "dd columns");
        }
        rp.items = cnt;
        if (tpc) {
            const int all[11] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
            DYNPY_CHECK(launch_rows2(rp, all, 11, false, n_sm, st), "dd rows");
        } else {
            const int plain[10] = {0, 1, 2, 3, 4, 6, 7, 8, 9, 10};
            const int gr[1] = {5};
            DYNPY_CHECK(launch_rows2(rp, plain, 10, false, n_sm, st), "dd rows");
            DYNPY_CHECK(launch_rows2(rp, gr, 1, true, n_sm, st), "dd rows (G_r)");
        }
    }
    return DYNPY_OK;
}

}  // namespace dynpy
