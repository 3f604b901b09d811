// Host-side driver of the FFT engine: picks the plan for a transform length M, sizes the
// grids in multiples of the SM count, and loops batches through the T workspace.
#pragma once
#include "fft_engine.cuh"
#include "fft_rows2.cuh"
#include "common.h"
#include <stdlib.h>
#include <string.h>

namespace dynpy {

// defined in dd2.cu
cudaError_t launch_rows2(Rows2Params p, const int* sel, int nsel, bool fix, int n_sm, cudaStream_t st);
cudaError_t launch_rows2_write(Rows2Params p, int mode, int n_sm, cudaStream_t st);
cudaError_t launch_rows_to_natural(const cd* T, double* out, int M1, int rs, i64 n_fft, i64 f_base, i64 n_out, i64 n_series,
                                   int planes, double scale, cudaStream_t st);

template <int L, bool TWO_PASS, class Loader, int MODE>
static cudaError_t launch_rows_L(const Loader& ld, const EpiParams& ep, const cd* tw, i64 f0, i64 f1, int gy,
                                 int M1, cudaStream_t st) {
    auto kern = k_rows<L, TWO_PASS, Loader, MODE>;
    constexpr int NT = (L / Plan<L>::RL) < 32 ? 32 : (L / Plan<L>::RL);
    size_t smem = (size_t)pad16(L) * sizeof(cd);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(TWO_PASS ? M1 : 1, gy);
    ProfScope prof(PROF_ROWS, st);
    kern<<<grid, NT, smem, st>>>(ld, ep, tw, f0, f1, gy, M1);
    return cudaGetLastError();
}

template <class Loader, int MODE>
static cudaError_t launch_rows_single(i64 M, const Loader& ld, const EpiParams& ep, const cd* tw, i64 f0, i64 f1,
                                      int gy, cudaStream_t st) {
    switch (M) {
        case 32: return launch_rows_L<32, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 64: return launch_rows_L<64, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 128: return launch_rows_L<128, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 256: return launch_rows_L<256, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 512: return launch_rows_L<512, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 1024: return launch_rows_L<1024, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 2048: return launch_rows_L<2048, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        case 4096: return launch_rows_L<4096, false, Loader, MODE>(ld, ep, tw, f0, f1, gy, 1, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int M1, int C, class Loader>
static cudaError_t launch_cols_M1(const Loader& ld, cd* T, const cd* tw, i64 M, i64 f0, i64 f1, int gy,
                                  cudaStream_t st) {
    auto kern = k_cols<M1, C, Loader>;
    constexpr int NT = (M1 / Plan<M1>::RL) * C;
    size_t smem = (Plan<M1>::S >= 2) ? (size_t)pad16(M1) * C * sizeof(cd) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(DYNPY_ROW_LEN / C, gy);
    ProfScope prof(PROF_COLS, st);
    kern<<<grid, NT, smem, st>>>(ld, T, tw, M, 1.0 / (double)M, f0, f1, gy);
    return cudaGetLastError();
}

template <int M1, class Loader>
static cudaError_t launch_cols_tpcg(const Loader& ld, cd* T, i64 M, i64 f0, i64 f1, int gy, cudaStream_t st) {
    const size_t smem = Loader::DENSE ? 0 : (size_t)(M1 / 2) * 256 * sizeof(cd);  // per-thread sample stash
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_cols_tpcg<M1, Loader>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    ProfScope prof(PROF_COLS, st);
    k_cols_tpcg<M1, Loader><<<dim3(DYNPY_ROW_LEN / 256, gy), 256, smem, st>>>(ld, T, 1.0 / (double)M, f0, f1, gy);
    return cudaGetLastError();
}

template <class Loader>
static cudaError_t launch_cols(int M1, const Loader& ld, cd* T, const cd* tw, i64 M, i64 f0, i64 f1, int gy,
                               cudaStream_t st) {
    switch (M1) {
        case 50: return launch_cols_tpcg<50, Loader>(ld, T, M, f0, f1, gy, st);
        case 20: return launch_cols_tpcg<20, Loader>(ld, T, M, f0, f1, gy, st);
        case 24: return launch_cols_tpcg<24, Loader>(ld, T, M, f0, f1, gy, st);
        case 40: return launch_cols_tpcg<40, Loader>(ld, T, M, f0, f1, gy, st);
        case 2: return launch_cols_M1<2, 256, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 4: return launch_cols_M1<4, 256, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 8: return launch_cols_M1<8, 256, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 16: return launch_cols_tpcg<16, Loader>(ld, T, M, f0, f1, gy, st);
        case 32: return launch_cols_tpcg<32, Loader>(ld, T, M, f0, f1, gy, st);
        case 64: return launch_cols_M1<64, 64, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 128: return launch_cols_M1<128, 32, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 256: return launch_cols_M1<256, 16, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 512: return launch_cols_M1<512, 8, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 1024: return launch_cols_M1<1024, 4, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 2048: return launch_cols_M1<2048, 2, Loader>(ld, T, tw, M, f0, f1, gy, st);
        case 4096: return launch_cols_M1<4096, 1, Loader>(ld, T, tw, M, f0, f1, gy, st);
        default: return cudaErrorInvalidValue;
    }
}

static inline int ceil_div_i(i64 a, i64 b) { return (int)((a + b - 1) / b); }

// Forward transforms of FFTs [0, n_fft) described by `ld`, consumed by epilogue MODE.
//   M        transform length (power of two, 32 <= M <= 2^24)
//   nslots   FFTs per item (f % nslots selects the accumulator in ACC mode); batches are whole items
//   T, T_cap workspace for the two-pass split: T_cap = number of FFTs it can hold (M*16 bytes each)
template <class Loader, int MODE>
static cudaError_t run_forward(const Loader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T,
                               i64 T_cap, const cd* tw, int n_sm, cudaStream_t st) {
    if (n_fft <= 0) return cudaSuccess;
    if (!fft_length_ok(M)) return cudaErrorInvalidValue;
    const int target = 4 * n_sm;  // CTAs in flight we aim for (>= 2 waves at 2 CTA/SM)
    if (M <= DYNPY_ROW_LEN) {
        i64 items = (n_fft + nslots - 1) / nslots;
        int ys = (int)(items < target ? items : target);
        if (ys < 1) ys = 1;
        return launch_rows_single<Loader, MODE>(M, ld, ep, tw, 0, n_fft, ys * nslots, st);
    }
    const int M1 = (int)(M / DYNPY_ROW_LEN);
    i64 cap_items = T_cap / nslots;
    if (!T || cap_items < 1) return cudaErrorMemoryAllocation;
    i64 items = (n_fft + nslots - 1) / nslots;
    for (i64 it0 = 0; it0 < items; it0 += cap_items) {
        i64 it1 = it0 + cap_items < items ? it0 + cap_items : items;
        i64 f0 = it0 * nslots, f1 = it1 * nslots < n_fft ? it1 * nslots : n_fft;
        i64 nb = f1 - f0;
        // first pass: (4096/C) tiles x gy CTAs
        int tiles = M1 <= 50 ? DYNPY_ROW_LEN / 256 : M1;  // == 4096 / (columns per CTA)
        int gy1 = ceil_div_i(target, tiles);
        if (gy1 > nb) gy1 = (int)nb;
        if (gy1 < 1) gy1 = 1;
        cudaError_t e = launch_cols<Loader>(M1, ld, T, tw, M, f0, f1, gy1, st);
        if (e != cudaSuccess) return e;
        // second pass: M1 rows x (nslots * ys) CTAs
        int ys = ceil_div_i(target, (i64)M1 * nslots);
        if (ys > it1 - it0) ys = (int)(it1 - it0);
        if (ys < 1) ys = 1;
        if constexpr (MODE == EPI_ACC_POWER) {
            // accumulate-power rows: the balanced register-accumulating kernel of fft_rows2.cuh
            if (nb % nslots != 0) return cudaErrorInvalidValue;
            Rows2Params rp;
            memset(&rp, 0, sizeof(rp));
            rp.T = T; rp.S = ep.out; rp.M = M; rp.M1 = M1; rp.rs = DYNPY_ROW_LEN; rp.nslots = nslots; rp.items = nb / nslots;
            rp.acc_pack = ep.acc_pack; rp.twp = tw;
            int sel[16];
            for (int s = 0; s < nslots; ++s) sel[s] = s;
            e = launch_rows2(rp, sel, nslots, false, n_sm, st);
        } else {
            // write epilogues (per-series power spectra, inverse transforms): the same row kernel.  Real outputs
            // (inverse transforms) are written back into T row by row and transposed to natural order by a second
            // kernel.  For A/B, DYNPY_ROWS_GENERIC=1 keeps the plain radix kernel k_rows for all epilogues, =2 for the
            // real-output modes only, =3 uses the row kernel's scattered 8-byte stores for them
            static int generic = -1;
            if (generic < 0) { const char* g = getenv("DYNPY_ROWS_GENERIC"); generic = g ? atoi(g) : 0; }
            if (generic == 0 || generic == 3 || (generic == 2 && MODE == EPI_WRITE_POWER)) {
                Rows2Params rp;
                memset(&rp, 0, sizeof(rp));
                rp.T = T; rp.M = M; rp.M1 = M1; rp.rs = DYNPY_ROW_LEN; rp.items = nb; rp.twp = tw;
                rp.out = ep.out; rp.f_base = f0; rp.n_out = ep.n_out; rp.n_series = ep.n_series; rp.scale = ep.scale;
                if (MODE != EPI_WRITE_POWER && generic == 0) {
                    e = launch_rows2_write(rp, ROWS2_WRITE_INPLACE, n_sm, st);
                    if (e == cudaSuccess)
                        e = launch_rows_to_natural(T, ep.out, M1, DYNPY_ROW_LEN, nb, f0, ep.n_out, ep.n_series,
                                                   MODE == EPI_WRITE_REAL2 ? 2 : 1, ep.scale, st);
                } else {
                    e = launch_rows2_write(rp, MODE, n_sm, st);
                }
            } else {
                TLoader tl;
                tl.T = T; tl.M1 = M1; tl.f_base = f0;
                e = launch_rows_L<DYNPY_ROW_LEN, true, TLoader, MODE>(tl, ep, tw, f0, f1, ys * nslots, M1, st);
            }
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace dynpy
