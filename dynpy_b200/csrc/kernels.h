// Launch wrappers of kernels_misc.cu and the engine entry points of engine_*.cu.
#pragma once
#include "common.h"
#include "fft_engine.cuh"

namespace dynpy {

// Dipolar FFT slots per pair-pair (p, q)  ->  accumulator (CSV column order of avg-acf.csv:
// 0 G_Y20, 1 G_Y21, 2 G_Y22, 3 G_r, 4 G_20, 5 G_21, 6 G_22):
//   0: Y20_p + i Y20_q -> 0      1: Y21_p -> 1    2: Y21_q -> 1    3: Y22_p -> 2   4: Y22_q -> 2
//   5: r3_p + i r3_q (mean removed) -> 3          6: F20_p + i F20_q -> 4
//   7: F21_p -> 5   8: F21_q -> 5   9: F22_p -> 6   10: F22_q -> 6
#define DD_NSLOTS 11
#define DD_NACC 7

cudaError_t launch_init_twiddles(double2* tw, cudaStream_t st);
cudaError_t launch_wrap(const double* x, double* out, i64 n, double L, int n_sm, cudaStream_t st);
cudaError_t launch_pdist(const double* ux, const double* uy, const double* uz, i64 stride, i64 n, double a,
                         double b, double c, const i64* index, double dmax, double* dx, double* dy, double* dz,
                         double* dr, i64* a0, i64* a1, i64* prj, i64* count, i64* block_counts, cudaStream_t st);
cudaError_t launch_supercell(const double* pos, i64 n, i64 F, i64 frame, double a, double* out, cudaStream_t st);
cudaError_t launch_dd_words(const double* X, i64 A, i64 N, const int* pi, const int* pj, i64 pair0, i64 n_pairs,
                            i64 items, double a, double b, double c, double* Z, double* bias, cudaStream_t st);
cudaError_t launch_cart_to_dipolar(const double* dx, const double* dy, const double* dz, const double* dr, i64 n,
                                   double* out, int n_sm, cudaStream_t st);
cudaError_t launch_dd_pair_count(const double* X, i64 N, const int* pi, const int* pj, i64 n_pairs, double a,
                                 double b, double c, double rmax, i64* counts, int* frame_hit, cudaStream_t st);
cudaError_t launch_topology_check(const double* X, i64 A, i64 N, i64 n0, i64 n1, const double* rcov, double extra,
                                  const unsigned char* expect, double a, double b, double c, double dmax,
                                  unsigned long long* out, i64* events, i64 cap, cudaStream_t st);
cudaError_t launch_dd_ragged_words(const double* X, i64 N, i64 NC, const int* pi, const int* pj, i64 pair0, i64 npairs,
                                   double a, double b, double c, double rmax, double* W, int* fidx, i64* lens,
                                   double* scale, double* bias, const unsigned char* mask, cudaStream_t st);
cudaError_t launch_dd_ragged_presence(const int* fidx, const i64* lens, i64 N, i64 npairs, i64* pair_frames,
                                      int* frame_hit, cudaStream_t st);
cudaError_t launch_dd_ragged_scatter(const double* acf, const int* fidx, const i64* lens, i64 N, i64 NC, i64 nfft,
                                     double* G, cudaStream_t st);
cudaError_t launch_sr_angmom(const double* pos, const double* vel, i64 N, const int* mol_atoms, i64 n_mol, int nat,
                             const double* mass, const int* axis_atoms, int axis_rule, double a, double b, double c,
                             double inv_dt, double* J, double* omega, double* axes, cudaStream_t st);
cudaError_t launch_spectrum_to_natural(const double* in, double* out, i64 n_spec, int M1, cudaStream_t st);
cudaError_t launch_simpson(const double* y, const double* x, i64 n, int n_cols, i64 col_stride, int mode,
                           double* out, cudaStream_t st);

// finite-frequency spectral densities (spectral.cu)
i64 dft_workspace_bytes(i64 n);
int dft_run(const double* re, const double* im, i64 n_series, i64 n, cd* out, i64 n_out, void* ws, i64 ws_bytes,
            int n_sm, cudaStream_t st);
cudaError_t launch_cutoff_bounds(const double* y, i64 n, int n_cols, i64 col_stride, int window, double tol,
                                 i64* bounds, int n_sm, cudaStream_t st);

// dipolar fast path (dd2.cu): warp-autonomous fused column kernel + row kernel v2
bool dd2_supported(i64 M);
i64 dd2_workspace_bytes(i64 M, i64 N, i64 items);
int dd2_run(const double* X, i64 N, const int* pi, const int* pj, i64 P, double a, double b, double c, double* S,
            i64 M, void* ws, i64 ws_bytes, const cd* tw, int n_sm, cudaStream_t st);

// engine entry points (one translation unit per loader so they compile in parallel)
cudaError_t forward_planar_acc(const PlanarLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T,
                               i64 T_cap, const cd* tw, int n_sm, cudaStream_t st);
cudaError_t forward_planar_power(const PlanarLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T,
                                 i64 T_cap, const cd* tw, int n_sm, cudaStream_t st);
cudaError_t forward_qr_acc(const QrLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T, i64 T_cap,
                           const cd* tw, int n_sm, cudaStream_t st);
cudaError_t forward_spectrum_real(const SpectrumLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots,
                                  cd* T, i64 T_cap, const cd* tw, int n_sm, cudaStream_t st);

cudaError_t forward_spectrum_real2(const SpectrumPairLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, cd* T,
                                   i64 T_cap, const cd* tw, int n_sm, cudaStream_t st);
cudaError_t launch_spectrum_even_natural(const double* in, double* out, i64 n_spec, i64 M, cudaStream_t st);
// rows of the packed even-spectra array for `pairs` ragged pairs (7 series each): 14 per two pairs, 8 for a last
// odd pair (one of them zero); always even
inline i64 dd_pack_rows(i64 pairs) { return 14 * (pairs / 2) + ((pairs & 1) ? 8 : 0); }

}  // namespace dynpy
