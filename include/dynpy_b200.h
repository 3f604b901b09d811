/* dynpy_b200 — C-ABI of the B200-native correlation-function hot path of DynPy.
 *
 * The reference (jautschbach/dynpy) is pure Python and has no FFI of its own; the
 * boundary is a handful of Python call sites.  Each entry point below
 * names the reference call site(s) it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - fp64 / int64 / int32 exactly as declared; no torch types cross this boundary;
 *   - `stream` is a cudaStream_t (0 = legacy default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise, re-entrant per stream;
 *   - return value: 0 ok, <0 error (DYNPY_ERR_*); dynpy_last_error() gives the text;
 *   - no C++ exception crosses the ABI; one process per GPU for multi-GPU use.
 *
 * Trajectory layout:  coords[atom][axis][frame]  (frame contiguous), bohr.
 * EFG layout:         efg[nucleus][comp][frame], comp = Vxx,Vyy,Vzz,Vxy,Vxz,Vyz.
 * Spectrum layout:    accumulated |X|^2 in the engine's internal order (only
 *                     dynpy_ifft_acf_f64 interprets it); several accumulators are
 *                     stored back to back, M doubles each.
 */
#ifndef DYNPY_B200_H
#define DYNPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNPY_OK 0
#define DYNPY_ERR_INVALID (-1)
#define DYNPY_ERR_LAUNCH (-2)
#define DYNPY_ERR_WORKSPACE (-3)
#define DYNPY_ERR_NO_DEVICE (-4)

#define DYNPY_SIMPSON_CARTWRIGHT 0 /* scipy >= 1.11 (even N: Cartwright end correction) */
#define DYNPY_SIMPSON_AVG 1        /* scipy < 1.11 even='avg' (reproduces the shipped golden file) */

#define DYNPY_AXIS_METHANE 0 /* z = d0/|d0|, x = z x d1, y = z x x   (SRrax.py:389-395, also methyl/ring) */
#define DYNPY_AXIS_WATER 1   /* Z = -(d0+d1)/|.|, X = Z x d0, Y = Z x X (SRrax.py:397-404) */

typedef void* dynpy_stream_t; /* cudaStream_t */

int dynpy_version(void);
const char* dynpy_last_error(void);
/* number of SMs of the current device (grids are sized in multiples of it); <0 on error */
int dynpy_device_sm_count(void);

/* Optional per-kernel-class device timing (CUDA events on the launching stream), for live
 * roofline numbers: classes 0 dipolar word generator, 1 FFT column pass, 2 FFT row pass, 3 other.
 * dynpy_profile_read synchronises the device, returns the summed milliseconds and launch counts
 * since the previous read (host arrays of 4) and resets them. */
int dynpy_profile_enable(int on);
int dynpy_profile_read(double* ms_host, int64_t* launches_host);

/* Padded transform length used for a series of n samples: the smallest of 2^k (k >= 5) and
 * 4096 x {20, 24, 40, 50} that is >= 2n-1 (e.g. 204800 for 81920 < n <= 102400: 22 % fewer points than
 * 2^18; mixed-radix column pass).  Any M >= 2n-1 gives the same linear ACF as the reference's n=2N
 * (qrax.py:153).  Every entry point accepts M = a power of two in [32, 2^24] or 4096 x {20, 24, 40, 50}. */
int64_t dynpy_fft_length(int64_t n);

/* Bytes of workspace that let `n_fft` transforms of length M be in flight in the two-pass
 * split (0 for M <= 4096). */
int64_t dynpy_fft_workspace_bytes(int64_t M, int64_t n_fft);

/* ---- Wiener–Khinchin, batched ---------------------------------------------------------
 * replaces  wiener_khinchin(f)  (qrax.py:147-156, ddrax.py:276-285, SRrax.py:216-225) applied
 * per series by correlate() (qrax.py:158-166, ddrax.py:287-295, SRrax.py:226-230), summed
 * over the series:   spectrum[k] += sum_s scale_s^2 |FFT_M(series_s)[k]|^2 .
 *   re, im          series s starts at re + s*series_stride (im may be NULL: real series)
 *   pack_real_pairs !=0 (im must be NULL): series 2t and 2t+1 share one complex transform
 *   scale           optional per-series multiplier applied before the transform (NULL: 1)
 *   spectrum        [M] accumulator (caller zeroes it), internal order
 */
int dynpy_wk_acf_sum_f64(const double* re, const double* im, int64_t n_series, int64_t n,
                         int64_t series_stride, int pack_real_pairs, const double* scale,
                         double* spectrum, int64_t M, void* ws, int64_t ws_bytes,
                         dynpy_stream_t stream);

/* acf[a][k] = scale * Re(FFT_M(spectrum_a))[k] , k < n_out  — the one inverse transform per
 * output column that replaces np.fft.ifft(...)[:N].real/N of every series (qrax.py:154-155);
 * pass scale = 1/(M*N) for the reference's biased estimator. */
int dynpy_ifft_acf_f64(const double* spectrum, int n_acc, int64_t M, double* acf, int64_t n_out,
                       double scale, void* ws, int64_t ws_bytes, dynpy_stream_t stream);

/* Per-series variant (no sum): acf[s][k] = wiener_khinchin(series_s)[k], k < n — what
 * correlate() returns per group when the individual ACFs are written out
 * (Jacfs_all.csv, SRparse.py:82,89). */
int64_t dynpy_wk_each_workspace_bytes(int64_t M, int64_t n_series);
int dynpy_wk_acf_each_f64(const double* re, const double* im, int64_t n_series, int64_t n,
                          int64_t series_stride, double* acf, int64_t M, void* ws, int64_t ws_bytes,
                          dynpy_stream_t stream);

/* ---- quadrupolar --------------------------------------------------------------------------
 * replaces cart_to_spatial + groupby('label').apply(correlate)  (qrax.py:59-60,115-166):
 * spectrum[3][M] accumulates, over the nuclei of this shard, the power spectra of
 * R2,0 (acc 0), R2,+-1 (acc 1), R2,+-2 (acc 2). */
int dynpy_qr_efg_to_spectrum_f64(const double* efg, int64_t n_nuclei, int64_t n_frames,
                                 double* spectrum, int64_t M, void* ws, int64_t ws_bytes,
                                 dynpy_stream_t stream);

/* ---- dipolar --------------------------------------------------------------------------------
 * np.mod(x, L) with numpy's sign convention (universe.py:296-308, distances.py:33-45). */
int dynpy_wrap_f64(const double* x, double* out, int64_t n, double L, dynpy_stream_t stream);

/* pdist_ortho (distances.py:155-242) for one frame, bit-exact: candidates visited in
 * (aa,bb,cc) order, strict '<' from dmax^2, no FMA contraction, compacted in row-major
 * upper-triangle order.  ux/uy/uz are read with element stride `stride`.
 * Outputs have capacity n(n-1)/2 — or pass all seven as NULL for a count-only call and size
 * them from *count (device int64), which receives the number of rows.
 * ws: at least dynpy_pdist_workspace_bytes(n). */
int64_t dynpy_pdist_workspace_bytes(int64_t n);
int dynpy_pdist_ortho_f64(const double* ux, const double* uy, const double* uz, int64_t stride,
                          int64_t n, double a, double b, double c, const int64_t* index,
                          double dmax, double* dx, double* dy, double* dz, double* dr,
                          int64_t* atom0, int64_t* atom1, int64_t* projection, int64_t* count,
                          void* ws, int64_t ws_bytes, dynpy_stream_t stream);

/* ddrax.cart_to_dipolar_from_df (ddrax.py:110-126) as a stand-alone call: n pair vectors (dx, dy, dz, dr) ->
 * out[11][n] = r^-3, Y20, Re Y21, Im Y21, Re Y22, Im Y22, F20, Re F21, Im F21, Re F22, Im F22 (Cartesian
 * closed forms of the angle formulas; the fused dipolar path forms the same words on the fly). */
int dynpy_cart_to_dipolar_f64(const double* dx, const double* dy, const double* dz, const double* dr,
                              int64_t n, double* out, dynpy_stream_t stream);

/* 3x3x3 block of periodic images of one frame (neighbors._worker, neighbors.py:160-199):
 * coords [n_atoms][3][n_frames] (raw) -> out[3][27*n_atoms]; super atom p*n_atoms + l is atom l
 * shifted by (i, j, k)*a with image number p = 9(i+1) + 3(j+1) + (k+1), i, j, k in {-1, 0, 1}. */
int dynpy_supercell_f64(const double* coords, int64_t n_atoms, int64_t n_frames, int64_t frame,
                        double a, double* out, dynpy_stream_t stream);

/* Static-topology check: the reference re-detects bonds (dr <= rcov0 + rcov1 + bond_extra among
 * pairs with dr < dmax, distances.py:244-257) and molecules in EVERY frame; the CUDA path
 * indexes molecules once.  This verifies the equivalence: expect[i*n_atoms + j] (i<j) is the bond
 * flag of the indexing frame; out[0] = mismatching pair-frames, out[1] = first such frame
 * (UINT64_MAX if none).  coords WRAPPED. */
int dynpy_topology_check_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                             const double* rcov, double bond_extra, const uint8_t* expect, double a,
                             double b, double c, double dmax, uint64_t* out, dynpy_stream_t stream);
/* Same scan, additionally listing the differences: events[k] = (frame, i, j) for every pair-frame whose
 * bond flag differs from `expect` (unordered; only the first `capacity` are stored), out[0] = mismatching
 * pair-frames, out[1] = first such frame, out[2] = number of events found (may exceed capacity).
 * Lets the host reproduce the reference's per-frame molecule detection (universe.py:411-451,
 * SRparse.py:171-176: molecules that are not intact in every frame are dropped) from a static topology
 * plus the few frames that deviate from it. */
int dynpy_topology_events_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                              const double* rcov, double bond_extra, const uint8_t* expect, double a,
                              double b, double c, double dmax, int64_t* events, int64_t capacity,
                              uint64_t* out, dynpy_stream_t stream);

/* The same scan over the frames [frame_begin, frame_end) only (events carry absolute frame indices): one process
 * per GPU scans its block of frames and the host merges the lists (the scan is O(atoms^2 x frames) and would
 * otherwise be repeated on every rank).  Reference: universe.py:411-451 over all frames. */
int dynpy_topology_events_range_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                                    int64_t frame_begin, int64_t frame_end, const double* rcov,
                                    double bond_extra, const uint8_t* expect, double a, double b, double c,
                                    double dmax, int64_t* events, int64_t capacity, uint64_t* out,
                                    dynpy_stream_t stream);

/* frames in which each pair's minimum-image distance is < rmax  (the presence rule of
 * compute_atom_two, DDparse.py:148): counts[p] in [0, n_frames]. coords are WRAPPED.
 * frame_hit (optional, [n_frames], caller-zeroed) is set to 1 where any pair is present. */
int dynpy_dd_pair_count_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                            const int32_t* pair_i, const int32_t* pair_j, int64_t n_pairs,
                            double a, double b, double c, double rmax, int64_t* counts,
                            int32_t* frame_hit, dynpy_stream_t stream);

/* Fused dipolar path for pairs present in every frame; replaces pdist_ortho per frame +
 * cart_to_dipolar_from_df + groupby(pair).apply(correlate) + the pair sum
 * (distances.py:131-145, ddrax.py:110-126,287-295, DDparse.py:81-104):
 * spectrum[7][M] accumulates sum_pairs |FFT|^2 of Y20,Y21,Y22,(r^-3 - mean),F20,F21,F22
 * (CSV column order of avg-acf.csv).  coords are WRAPPED, [n_atoms][3][n_frames].
 * ws: dynpy_dd_workspace_bytes(n_frames, M, items) for `items` pair-pairs per batch (>=1). */
int64_t dynpy_dd_workspace_bytes(int64_t n_frames, int64_t M, int64_t items);
int dynpy_dd_pairs_to_spectrum_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                                   const int32_t* pair_i, const int32_t* pair_j, int64_t n_pairs,
                                   double a, double b, double c, double* spectrum, int64_t M,
                                   void* ws, int64_t ws_bytes, dynpy_stream_t stream);

/* Ragged dipolar path (pairs absent from some frames, reference quirk B.3): every pair's
 * series is the compressed sequence of its present frames, normalised by its own length,
 * and lag k is binned at the k-th present frame:  G[7][n_frames] += ACF_p[k]/N_p at
 * frame_of(p,k).  ws: dynpy_dd_ragged_workspace_bytes(n_frames, M, pairs_per_batch). */
int64_t dynpy_dd_ragged_workspace_bytes(int64_t n_frames, int64_t M, int64_t pairs);
int dynpy_dd_pairs_ragged_to_acf_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                                     const int32_t* pair_i, const int32_t* pair_j, int64_t n_pairs,
                                     double a, double b, double c, double rmax, double* G,
                                     int64_t M, void* ws, int64_t ws_bytes, dynpy_stream_t stream);
/* The same with an explicit per-frame filter: mask[p*n_frames + f] != 0 where the pair ROW of frame f passes
 * the reference's molecule filter (DDparse.py:57-60: `intra` / `inter` compare the per-frame molecule ids, so
 * a pair changes status in frames where the bond list differs); NULL = no filter.  pair_frames[p] (optional)
 * receives the number of frames the pair is present in, frame_hit[f] (optional) is set to 1 for every frame
 * some pair is present in. */
int dynpy_dd_pairs_masked_to_acf_f64(const double* coords, int64_t n_atoms, int64_t n_frames,
                                     const int32_t* pair_i, const int32_t* pair_j, int64_t n_pairs,
                                     double a, double b, double c, double rmax, const uint8_t* mask,
                                     double* G, int64_t* pair_frames, int32_t* frame_hit, int64_t M,
                                     void* ws, int64_t ws_bytes, dynpy_stream_t stream);

/* The same call for pairs of one LENGTH CLASS: every pair of the call is present in at most n_cap frames
 * (n_cap <= n_frames; the caller knows the counts from dynpy_dd_pair_count_f64), so the compressed series, the
 * transform length M >= 2 n_cap - 1 and the workspace (dynpy_dd_ragged_workspace_bytes(n_cap, M, pairs)) are sized
 * by n_cap instead of n_frames: a pair that is inside rmax for 5 % of a long trajectory costs a 5 %-length
 * transform.  Pairs with more than n_cap present frames are truncated (caller error). */
int dynpy_dd_pairs_ragged_capped_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const int32_t* pair_i,
                                     const int32_t* pair_j, int64_t n_pairs, double a, double b, double c,
                                     double rmax, const uint8_t* mask, double* G, int64_t* pair_frames,
                                     int32_t* frame_hit, int64_t n_cap, int64_t M, void* ws, int64_t ws_bytes,
                                     dynpy_stream_t stream);

/* ---- spin-rotation -------------------------------------------------------------------------
 * replaces SR_func1 per (molecule, frame) (SRrax.py:438-510) and the backward-difference
 * velocities (SRparse.py:127-134).  pos is RAW [n_atoms][3][n_frames]; vel may be NULL
 * (v_t = (x_t - x_{t-1}) * inv_dt, outputs cover frames 1..n_frames-1) or explicit
 * [n_atoms][3][n_frames] (outputs cover all frames).  mol_atoms[m][k]: atom indices of
 * molecule m; mass[k] per atom of the molecule type; axis_atoms[m][4] = (a0,b0,a1,b1):
 * d0 = minimum-image vector of atoms (a0,b0) as pdist_ortho would emit it, d1 of (a1,b1).
 * J, omega: [n_mol][3][n_out]; axes: [n_mol][9][n_out] (xx,xy,xz,yx,...,zz); any may be NULL. */
int dynpy_sr_angmom_f64(const double* pos, const double* vel, int64_t n_atoms, int64_t n_frames,
                        const int32_t* mol_atoms, int64_t n_mol, int n_at_mol, const double* mass,
                        const int32_t* axis_atoms, int axis_rule, double a, double b, double c,
                        double inv_dt, double* J, double* omega, double* axes,
                        dynpy_stream_t stream);

/* ---- reductions -------------------------------------------------------------------------------
 * scipy.integrate.simpson(y_col, x=x) for n_cols columns (qrax.py:193-195, ddrax.py:208-210,
 * SRrax.py:576): y[col*col_stride + k], out[col]. */
int dynpy_simpson_f64(const double* y, const double* x, int64_t n, int n_cols, int64_t col_stride,
                      int mode, double* out, dynpy_stream_t stream);

/* ---- finite-frequency spectral densities ----------------------------------------------------------
 * numpy.fft.fft of ARBITRARY length n, as ddrax.fourier_T applies it to the G_2m columns of avg-acf
 * (ddrax.py:169-184): out[s][k] = sum_j (re[s][j] + i im[s][j]) exp(-2 pi i j k / n), k < n_out <= n;
 * im may be NULL (real series).  out is interleaved complex [n_series][n_out][2].
 * ws: dynpy_dft_workspace_bytes(n). */
int64_t dynpy_dft_workspace_bytes(int64_t n);
int dynpy_dft_f64(const double* re, const double* im, int64_t n_series, int64_t n, double* out,
                  int64_t n_out, void* ws, int64_t ws_bytes, dynpy_stream_t stream);
/* Integration bound of spec_dens_extr_narr / spectral_dens with cutoff=True (ddrax.py:188-198,
 * qrax.py:172-182): bounds[col] = first i < n-1 whose trailing `window`-point mean of y/y[0] is within
 * `tol` of zero (pandas rolling(window).mean() + numpy.isclose(., 0, atol=tol)), else n-1.
 * y[col*col_stride + k]; bounds: device int64 [n_cols]. */
int dynpy_cutoff_bounds_f64(const double* y, int64_t n, int n_cols, int64_t col_stride, int window,
                            double tol, int64_t* bounds, dynpy_stream_t stream);

/* ---- trajectory ingestion (HOST function; no GPU needed) -----------------------------------------
 * replaces the text handling of parseMD.parse_xyz / parse_tinker_md (parseMD.py:177-278):
 * blocks of  <header_lines> lines + nat atom lines  "[cols...] sym x y z [cols...]";
 *   first_col     column of the symbol (0: xyz, 1: Tinker .arc); x y z follow it (-1: no symbol, x y z first)
 *   printed_host  1-indexed printed frame numbers to keep, ascending  [n_printed]
 *   scale_div     every value is DIVIDED by it (0.529177: Angstrom -> bohr exactly like the reference)
 *   out_host      [n_printed][nat][3] fp64 (caller-owned, may be pinned memory)
 *   symbols_host  optional [nat][8] NUL-padded raw symbol tokens of the first kept frame
 *   n_threads     <= 0: all hardware threads
 * Fortran 'D' exponents are accepted; conversion is correctly rounded (== Python float()). */
int dynpy_parse_md_blocks_host(const char* path, int64_t nat, int32_t header_lines, int32_t first_col,
                               const int64_t* printed_host, int64_t n_printed, double scale_div,
                               double* out_host, char* symbols_host, int32_t n_threads);

/* Same reader for the QE .pos / .vel layout of parseMD.parse_qe_md (parseMD.py:67-153): one header line
 * "<step> <time>" per block, then nat lines "x y z" (first_col = -1: no symbol token; .vel lines carry one
 * leading token, first_col = 0).  header_first_host (optional, [n_printed]) receives the first token of
 * each kept block's header line as a number: the reference takes its frame ids from there
 * (parseMD.py:84-92).  Use scale_div = 1.0: the reference does not convert QE coordinates. */
int dynpy_parse_md_blocks_hdr_host(const char* path, int64_t nat, int32_t header_lines, int32_t first_col,
                                   const int64_t* printed_host, int64_t n_printed, double scale_div,
                                   double* out_host, char* symbols_host, double* header_first_host,
                                   int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* DYNPY_B200_H */
