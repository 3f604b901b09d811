// dynpy_b200 — non-FFT kernels of the hot path (sm_100a): coordinate wrap, bit-exact
// pdist_ortho, dipolar word generator, pair presence count, spin-rotation angular momentum,
// Simpson reduction.  Reference lines are cited per kernel.
#include <stdint.h>
#include "kernels.h"
#include "dd_device.cuh"

namespace dynpy {

// ------------------------------------------------------------------ twiddle table
__global__ void k_init_twiddles(double2* tw) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 4096) {
        double s, c;
        sincospi(2.0 * (double)j / 4096.0, &s, &c);
        tw[j] = make_double2(c, -s);
    }
}
cudaError_t launch_init_twiddles(double2* tw, cudaStream_t st) {
    k_init_twiddles<<<16, 256, 0, st>>>(tw);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ wrap
__global__ void k_wrap(const double* __restrict__ x, double* __restrict__ out, i64 n, double L) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = np_mod(x[i], L);
}
cudaError_t launch_wrap(const double* x, double* out, i64 n, double L, int n_sm, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)n_sm * 16) blocks = (i64)n_sm * 16;
    k_wrap<<<(unsigned)blocks, 256, 0, st>>>(x, out, n, L);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ pdist_ortho (bit-exact)
__device__ __forceinline__ void pair_from_linear(i64 p, i64 n, i64& i, i64& j) {
    // row-major upper triangle: row i starts at i*n - i*(i+1)/2
    double nn = (double)n;
    i64 ii = (i64)floor(((2.0 * nn - 1.0) - sqrt((2.0 * nn - 1.0) * (2.0 * nn - 1.0) - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    if (ii > n - 2) ii = n - 2;
    while (ii > 0 && ii * n - ii * (ii + 1) / 2 > p) --ii;
    while ((ii + 1) * n - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
    i = ii;
    j = p - (ii * n - ii * (ii + 1) / 2) + ii + 1;
}

template <bool WRITE>
__global__ void __launch_bounds__(256)
k_pdist(const double* __restrict__ ux, const double* __restrict__ uy, const double* __restrict__ uz, i64 stride,
        i64 n, double a, double b, double c, const i64* __restrict__ index, double dmax2, i64 npairs,
        i64* __restrict__ block_counts, double* dx, double* dy, double* dz, double* dr, i64* a0, i64* a1,
        i64* prj) {
    __shared__ int warp_cnt[8];
    const i64 p = (i64)blockIdx.x * 256 + threadIdx.x;
    MinImage r;
    r.found = false;
    i64 i = 0, j = 0;
    if (p < npairs) {
        pair_from_linear(p, n, i, j);
        r = min_image27(ux[i * stride], uy[i * stride], uz[i * stride], ux[j * stride], uy[j * stride],
                        uz[j * stride], a, b, c, dmax2);
    }
    const unsigned ball = __ballot_sync(0xffffffffu, r.found);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ball);
    __syncthreads();
    if (!WRITE) {
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_cnt[w];
            block_counts[blockIdx.x] = t;
        }
    } else {
        if (r.found) {
            i64 off = block_counts[blockIdx.x];  // exclusive prefix of this block
            for (int w = 0; w < warp; ++w) off += warp_cnt[w];
            off += __popc(ball & ((1u << lane) - 1u));
            dx[off] = r.dx; dy[off] = r.dy; dz[off] = r.dz;
            dr[off] = sqrt(r.r2);  // np.sqrt(dpr_) (distances.py:225): correctly rounded in both
            a0[off] = index[i]; a1[off] = index[j]; prj[off] = r.prj;
        }
    }
}

// neighbors._worker (neighbors.py:160-199): the 3x3x3 block of periodic images of one frame, image
// p = 9 (i+1) + 3 (j+1) + (k+1), i, j, k in {-1, 0, 1}; super atom p * n + l sits at x_l + i*a (etc.)
__global__ void k_supercell(const double* __restrict__ pos, i64 n, i64 F, i64 frame, double a, double* __restrict__ out) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;  // over 27 n
    if (t >= 27 * n) return;
    const int p = (int)(t / n);
    const i64 l = t - (i64)p * n;
    const int sh[3] = {p / 9 - 1, (p / 3) % 3 - 1, p % 3 - 1};
#pragma unroll
    for (int d = 0; d < 3; ++d) out[(i64)d * 27 * n + t] = __dadd_rn(pos[(l * 3 + d) * F + frame], __dmul_rn((double)sh[d], a));
}
cudaError_t launch_supercell(const double* pos, i64 n, i64 F, i64 frame, double a, double* out, cudaStream_t st) {
    k_supercell<<<(unsigned)((27 * n + 255) / 256), 256, 0, st>>>(pos, n, F, frame, a, out);
    return cudaGetLastError();
}

// exclusive scan of block counts by one CTA; total -> *count
__global__ void k_scan_counts(i64* __restrict__ v, i64 n, i64* __restrict__ count) {
    __shared__ i64 sh[1024];
    __shared__ i64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (i64 base = 0; base < n; base += 1024) {
        i64 idx = base + threadIdx.x;
        i64 x = idx < n ? v[idx] : 0;
        sh[threadIdx.x] = x;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            i64 t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        i64 incl = sh[threadIdx.x];
        if (idx < n) v[idx] = carry + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

cudaError_t launch_pdist(const double* ux, const double* uy, const double* uz, i64 stride, i64 n, double a,
                         double b, double c, const i64* index, double dmax, double* dx, double* dy, double* dz,
                         double* dr, i64* a0, i64* a1, i64* prj, i64* count, i64* block_counts,
                         cudaStream_t st) {
    const i64 npairs = n * (n - 1) / 2;
    if (npairs <= 0) return cudaMemsetAsync(count, 0, sizeof(i64), st);
    const i64 nblocks = (npairs + 255) / 256;
    const double dmax2 = dmax * dmax;  // dmax**2 (distances.py:177)
    k_pdist<false><<<(unsigned)nblocks, 256, 0, st>>>(ux, uy, uz, stride, n, a, b, c, index, dmax2, npairs,
                                                     block_counts, dx, dy, dz, dr, a0, a1, prj);
    k_scan_counts<<<1, 1024, 0, st>>>(block_counts, nblocks, count);
    if (!dx) return cudaGetLastError();  // count-only call: the caller sizes the outputs from *count
    k_pdist<true><<<(unsigned)nblocks, 256, 0, st>>>(ux, uy, uz, stride, n, a, b, c, index, dmax2, npairs,
                                                    block_counts, dx, dy, dz, dr, a0, a1, prj);
    return cudaGetLastError();
}

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;  // valid in thread 0
}

// grid (ceil(N/256), items): thread = one frame of pair-pair `item` (pairs 2*item, 2*item+1).
// Z[item][slot][2][N] planar; slots documented in kernels.h (DD_NSLOTS = 11).
__global__ void __launch_bounds__(256)
k_dd_words(const double* __restrict__ X, i64 A, i64 N, const int* __restrict__ pi, const int* __restrict__ pj,
           i64 pair0, i64 n_pairs, double a, double b, double c, double* __restrict__ Z,
           double* __restrict__ bias) {
    __shared__ double sh[8];
    const i64 item = blockIdx.y;
    const i64 n = (i64)blockIdx.x * 256 + threadIdx.x;
    double* z = Z + item * (i64)(DD_NSLOTS * 2) * N;
    double rsum[2] = {0.0, 0.0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const i64 p = pair0 + 2 * item + h;
        DDWords w;
        bool ok = (p < n_pairs) && (n < N);
        if (ok) {
            const i64 i = pi[p], j = pj[p];
            const double* xi = X + i * 3 * N;
            const double* xj = X + j * 3 * N;
            double dx, dy, dz, sx, sy, sz;
            axis_min(__ldg(xi + n), __ldg(xj + n), a, dx, sx);
            axis_min(__ldg(xi + N + n), __ldg(xj + N + n), b, dy, sy);
            axis_min(__ldg(xi + 2 * N + n), __ldg(xj + 2 * N + n), c, dz, sz);
            const double r2 = __dadd_rn(__dadd_rn(sx, sy), sz);
            w = dd_words(dx, dy, dz, r2);
            rsum[h] = w.r3;
        } else {
            w.y20 = w.y21r = w.y21i = w.y22r = w.y22i = w.r3 = w.f20 = w.f21r = w.f21i = w.f22r = w.f22i = 0.0;
        }
        if (n < N) {
            // packed real series: plane h of slots 0 (Y20), 5 (r^-3), 6 (F20)
            z[(0 * 2 + h) * N + n] = w.y20;
            z[(5 * 2 + h) * N + n] = w.r3;
            z[(6 * 2 + h) * N + n] = w.f20;
            // complex series of pair h: slots 1+h (Y21), 3+h (Y22), 7+h (F21), 9+h (F22)
            z[((1 + h) * 2 + 0) * N + n] = w.y21r; z[((1 + h) * 2 + 1) * N + n] = w.y21i;
            z[((3 + h) * 2 + 0) * N + n] = w.y22r; z[((3 + h) * 2 + 1) * N + n] = w.y22i;
            z[((7 + h) * 2 + 0) * N + n] = w.f21r; z[((7 + h) * 2 + 1) * N + n] = w.f21i;
            z[((9 + h) * 2 + 0) * N + n] = w.f22r; z[((9 + h) * 2 + 1) * N + n] = w.f22i;
        }
    }
    const double s0 = block_sum_256(rsum[0], sh);
    const double s1 = block_sum_256(rsum[1], sh);
    if (threadIdx.x == 0) {
        atomicAdd(bias + 2 * item, s0);
        atomicAdd(bias + 2 * item + 1, s1);
    }
}
cudaError_t launch_dd_words(const double* X, i64 A, i64 N, const int* pi, const int* pj, i64 pair0, i64 n_pairs,
                            i64 items, double a, double b, double c, double* Z, double* bias, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(bias, 0, sizeof(double) * 2 * items, st);
    if (e != cudaSuccess) return e;
    for (i64 it0 = 0; it0 < items; it0 += 32768) {
        i64 cnt = items - it0 < 32768 ? items - it0 : 32768;
        dim3 grid((unsigned)((N + 255) / 256), (unsigned)cnt);
        k_dd_words<<<grid, 256, 0, st>>>(X, A, N, pi, pj, pair0 + 2 * it0, n_pairs, a, b, c,
                                         Z + it0 * (i64)(DD_NSLOTS * 2) * N, bias + 2 * it0);
    }
    return cudaGetLastError();
}

// frames with r2 < rmax^2 per pair: one CTA per pair
__global__ void __launch_bounds__(256)
k_dd_pair_count(const double* __restrict__ X, i64 N, const int* __restrict__ pi, const int* __restrict__ pj,
                double a, double b, double c, double rmax2, i64* __restrict__ counts, int* __restrict__ frame_hit) {
    __shared__ double sh[8];
    const i64 p = blockIdx.x;
    const double* xi = X + (i64)pi[p] * 3 * N;
    const double* xj = X + (i64)pj[p] * 3 * N;
    double cnt = 0.0;
    for (i64 n = threadIdx.x; n < N; n += 256) {
        double d, sx, sy, sz;
        axis_min(__ldg(xi + n), __ldg(xj + n), a, d, sx);
        axis_min(__ldg(xi + N + n), __ldg(xj + N + n), b, d, sy);
        axis_min(__ldg(xi + 2 * N + n), __ldg(xj + 2 * N + n), c, d, sz);
        if (__dadd_rn(__dadd_rn(sx, sy), sz) < rmax2) {
            cnt += 1.0;
            if (frame_hit) frame_hit[n] = 1;  // every writer stores the same value
        }
    }
    const double t = block_sum_256(cnt, sh);
    if (threadIdx.x == 0) counts[p] = (i64)t;
}
// The same for N % 4 == 0 (every coordinate stream then starts on a 32-byte boundary): four consecutive frames per
// thread and iteration, the six 32-byte loads of an iteration all issued before the first use (the scalar form waits
// on six dependent 8-byte loads per frame: 11 long-scoreboard stall cycles per issue)
__device__ __forceinline__ void ldg256_nc(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
// grid = (pairs, chunks of PC_CHUNK frames), pairs fastest: all pairs sweep one chunk of frames while its coordinates
// (A x 3 x 4096 frames = 50 MB at 512 atoms) are in L2 — one CTA per pair streaming the whole trajectory re-read the
// second atom of every pair from DRAM (130 816 pairs x 2.4 MB = 314 GB at cfg2).  counts[] is zeroed by the launcher.
#define PC_CHUNK 4096
__global__ void __launch_bounds__(256)
k_dd_pair_count4(const double* __restrict__ X, i64 N, const int* __restrict__ pi, const int* __restrict__ pj,
                 double a, double b, double c, double rmax2, unsigned long long* __restrict__ counts,
                 int* __restrict__ frame_hit) {
    __shared__ double sh[8];
    const i64 p = blockIdx.x;
    const double* xi = X + (i64)pi[p] * 3 * N;
    const double* xj = X + (i64)pj[p] * 3 * N;
    const double L[3] = {a, b, c};
    const i64 n_begin = (i64)blockIdx.y * PC_CHUNK;
    const i64 n_end = n_begin + PC_CHUNK < N ? n_begin + PC_CHUNK : N;
    double cnt = 0.0;
    for (i64 n = n_begin + (i64)threadIdx.x * 4; n < n_end; n += 1024) {
        double ci[3][4], cj[3][4];
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            ldg256_nc(xi + ax * N + n, ci[ax]);
            ldg256_nc(xj + ax * N + n, cj[ax]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double d, s2[3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) axis_min(ci[ax][q], cj[ax][q], L[ax], d, s2[ax]);
            if (__dadd_rn(__dadd_rn(s2[0], s2[1]), s2[2]) < rmax2) {
                cnt += 1.0;
                if (frame_hit) frame_hit[n + q] = 1;  // every writer stores the same value
            }
        }
    }
    const double t = block_sum_256(cnt, sh);
    if (threadIdx.x == 0 && t > 0.0) atomicAdd(counts + p, (unsigned long long)t);
}
cudaError_t launch_dd_pair_count(const double* X, i64 N, const int* pi, const int* pj, i64 n_pairs, double a,
                                 double b, double c, double rmax, i64* counts, int* frame_hit, cudaStream_t st) {
    if (n_pairs <= 0) return cudaSuccess;
    const i64 chunks = (N + PC_CHUNK - 1) / PC_CHUNK;
    if ((N & 3) == 0 && ((uintptr_t)X & 31) == 0 && chunks <= 65535) {
        cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(i64) * n_pairs, st);
        if (e != cudaSuccess) return e;
        k_dd_pair_count4<<<dim3((unsigned)n_pairs, (unsigned)chunks), 256, 0, st>>>(
            X, N, pi, pj, a, b, c, rmax * rmax, reinterpret_cast<unsigned long long*>(counts), frame_hit);
    } else {
        k_dd_pair_count<<<(unsigned)n_pairs, 256, 0, st>>>(X, N, pi, pj, a, b, c, rmax * rmax, counts, frame_hit);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ static-topology check
// For every frame and every atom pair i<j: is (min-image dr <= rcov_i + rcov_j + extra) equal to
// the bond flag of the reference frame?  (The reference re-detects bonds/molecules in every
// frame, distances.py:244-257 + universe.py:411-451; the CUDA path indexes molecules once and
// verifies here that this is equivalent.)  out[0] = number of mismatching pair-frames,
// out[1] = first frame with a mismatch, out[2] = number of events listed.
//
// O(A^2 F) pair-frames, but almost all of them are far apart: a thread owns one frame and a tile of
// TOPO_TI atoms i whose coordinates stay in (float) registers; for every j it loads x_j, y_j, z_j once (coalesced
// 1 KB runs per warp: consecutive threads = consecutive frames) and screens the TI pairs on all three axes —
// nearest-image |d| against T = the largest bond threshold, with a margin that keeps the fp32 screen a superset.
// Only pairs inside the T-cube (a fraction (2T/L)^3) take the exact path: the bit-exact per-axis minimum image of
// the pair kernels and the reference's comparison (lanes of a warp are consecutive frames of the SAME pair, so the
// branch is nearly uniform).  The expected bonds of the tile's atoms are listed in shared memory once per CTA and
// verified in every frame first, so the screened sweep only has to find bonds that are NOT expected.  Coordinates
// are swept A / TOPO_TI times through L2 (the CTAs of one frame chunk are adjacent in the grid) instead of A / 2 times.
#define TOPO_TI 8
#define TOPO_NT 128
#define TOPO_BCAP 128
__device__ __forceinline__ bool topo_bond_exact(const double* __restrict__ X, i64 N, i64 n, i64 i, i64 j, double ri,
                                                double rj, double extra, double a, double b, double c, double dmax2) {
    double d, sx, sy, sz;
    axis_min(X[(i * 3 + 0) * N + n], __ldg(X + (j * 3 + 0) * N + n), a, d, sx);
    axis_min(X[(i * 3 + 1) * N + n], __ldg(X + (j * 3 + 1) * N + n), b, d, sy);
    axis_min(X[(i * 3 + 2) * N + n], __ldg(X + (j * 3 + 2) * N + n), c, d, sz);
    const double r2 = __dadd_rn(__dadd_rn(sx, sy), sz);
    return (r2 < dmax2) && (sqrt(r2) <= __dadd_rn(__dadd_rn(ri, rj), extra));
}
// Screen of one axis: is the nearest of the images -1, 0, +1 possibly within T?  min(|d|, ||d| - a|) <= T  is implied
// by  ||d| - a/2| >= a/2 - T  (equivalent for |d| <= a, a superset beyond): three instructions per axis, no fmin.
__global__ void __launch_bounds__(TOPO_NT)
k_topology_check(const double* __restrict__ X, i64 A, i64 N, const double* __restrict__ rcov, double extra,
                 const unsigned char* __restrict__ expect, double a, double b, double c, double dmax2,
                 double dmax, i64 tiles, unsigned long long* __restrict__ out,
                 i64* __restrict__ events, i64 cap, i64 n0, i64 n1) {
    __shared__ int s_nb;
    __shared__ double s_rmax[TOPO_NT / 32];
    __shared__ unsigned s_bond[TOPO_BCAP];   // (k << 24) | j : expected bonds (i0 + k, j), j > i0 + k
    // 1-D grid, atom tiles fastest: the CTAs of one frame chunk are adjacent and share its coordinates in L2
    const i64 chunk = blockIdx.x / tiles;
    const i64 n = n0 + chunk * TOPO_NT + threadIdx.x;   // frames [n0, n1) of the N the arrays hold
    const i64 i0 = (blockIdx.x - chunk * tiles) * TOPO_TI;
    const int ti = (int)(A - i0 < TOPO_TI ? A - i0 : TOPO_TI);
    // screening radius: the largest bond threshold any pair can have (2 max rcov + extra, never more than dmax), with
    // a margin that covers the rounding of the screen's own subtractions; formed per CTA (A loads, against the
    // ~A * TOPO_TI * TOPO_NT pair-frames the CTA screens)
    double rm = 0.0;
    for (i64 k = threadIdx.x; k < A; k += TOPO_NT) rm = fmax(rm, __ldg(rcov + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, o));
    if ((threadIdx.x & 31) == 0) s_rmax[threadIdx.x >> 5] = rm;
    if (threadIdx.x == 0) s_nb = 0;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < TOPO_NT / 32; ++w) rm = fmax(rm, s_rmax[w]);
    double T = fmin(2.0 * rm + extra, dmax);
    if (!(T > 0.0)) T = 0.0;
    T = T * (1.0 + 1e-9) + 1e-9;
    const double ha = 0.5 * a, hb = 0.5 * b, hc = 0.5 * c;
    const double ta = ha - T, tb = hb - T, tc = hc - T;
    for (int k = 0; k < ti; ++k) {
        const unsigned char* row = expect + (i0 + k) * A;
        for (i64 j = i0 + k + 1 + threadIdx.x; j < A; j += TOPO_NT)
            if (row[j]) {
                const int p = atomicAdd(&s_nb, 1);
                if (p < TOPO_BCAP) s_bond[p] = ((unsigned)k << 24) | (unsigned)j;
            }
    }
    __syncthreads();
    const int nb = s_nb;
    const bool listed = nb <= TOPO_BCAP && A < (1 << 24);   // else: expected flags are read from global memory
    if (n >= n1) return;
    unsigned long long bad = 0;
    auto report = [&](i64 i, i64 j) {
        ++bad;
        if (events) {  // (frame, i, j) of every bond flag that differs from the indexing frame's
            const unsigned long long k = atomicAdd(out + 2, 1ull);
            if ((i64)k < cap) { events[3 * k] = n; events[3 * k + 1] = i; events[3 * k + 2] = j; }
        }
    };
    // ---- expected bonds must be bonds in this frame
    if (listed) {
        for (int q = 0; q < nb; ++q) {
            const i64 i = i0 + (s_bond[q] >> 24), j = s_bond[q] & 0xffffffu;
            if (!topo_bond_exact(X, N, n, i, j, rcov[i], rcov[j], extra, a, b, c, dmax2)) report(i, j);
        }
    } else {
        for (int k = 0; k < ti; ++k)
            for (i64 j = i0 + k + 1; j < A; ++j)
                if (expect[(i0 + k) * A + j] &&
                    !topo_bond_exact(X, N, n, i0 + k, j, rcov[i0 + k], rcov[j], extra, a, b, c, dmax2))
                    report(i0 + k, j);
    }
    // ---- screened sweep: bonds that are not expected.  All three axes are screened at once, in fp32: the tile's
    // coordinates sit in 24 float registers, x_j / y_j / z_j are loaded (coalesced over frames) and converted once per
    // j, and a pair costs nine full-rate fp32 instructions; the thresholds are widened by far more than the rounding
    // of the conversions and differences (1e-5 of the box edge against ~2e-7), so the screen stays a superset and
    // every pair that passes (a fraction (2T/L)^3 ~ 0.6 % at cfg2) takes the exact fp64 path.  Screening x alone in
    // fp64 sent 80 % of the j iterations of an 8-atom tile into the slow branch: 59 ms per 768 atoms x 1e5 frames.
    float xi[TOPO_TI], yi[TOPO_TI], zi[TOPO_TI];
#pragma unroll
    for (int k = 0; k < TOPO_TI; ++k) {
        xi[k] = k < ti ? (float)X[((i0 + k) * 3 + 0) * N + n] : 0.0f;
        yi[k] = k < ti ? (float)X[((i0 + k) * 3 + 1) * N + n] : 0.0f;
        zi[k] = k < ti ? (float)X[((i0 + k) * 3 + 2) * N + n] : 0.0f;
    }
    const float haf = (float)ha, hbf = (float)hb, hcf = (float)hc;
    const float taf = (float)ta - (float)(1e-5 * a + 1e-6), tbf = (float)tb - (float)(1e-5 * b + 1e-6),
                tcf = (float)tc - (float)(1e-5 * c + 1e-6);
    for (i64 j = i0 + 1; j < A; ++j) {
        const float xj = (float)__ldg(X + (j * 3 + 0) * N + n);
        const float yj = (float)__ldg(X + (j * 3 + 1) * N + n);
        const float zj = (float)__ldg(X + (j * 3 + 2) * N + n);
        unsigned near = 0;
#pragma unroll
        for (int k = 0; k < TOPO_TI; ++k) {
            const bool nr = fabsf(fabsf(xi[k] - xj) - haf) >= taf && fabsf(fabsf(yi[k] - yj) - hbf) >= tbf &&
                            fabsf(fabsf(zi[k] - zj) - hcf) >= tcf;
            near |= nr ? (1u << k) : 0u;
        }
        if (j - i0 < TOPO_TI) near &= (1u << (j - i0)) - 1u;   // pairs with i0 + k < j only
        near &= (1u << ti) - 1u;
        while (near) {
            const int k = __ffs(near) - 1;
            near &= near - 1;
            const i64 i = i0 + k;
            if (topo_bond_exact(X, N, n, i, j, rcov[i], rcov[j], extra, a, b, c, dmax2) && !expect[i * A + j])
                report(i, j);
        }
    }
    if (bad) {
        atomicAdd(out, bad);
        atomicMin(out + 1, (unsigned long long)n);
    }
}
cudaError_t launch_topology_check(const double* X, i64 A, i64 N, i64 n0, i64 n1, const double* rcov, double extra,
                                  const unsigned char* expect, double a, double b, double c, double dmax,
                                  unsigned long long* out, i64* events, i64 cap, cudaStream_t st) {
    if (A < 2 || N < 1 || n1 <= n0) return cudaSuccess;
    const i64 tiles = (A + TOPO_TI - 1) / TOPO_TI, chunks = (n1 - n0 + TOPO_NT - 1) / TOPO_NT;
    if (tiles * chunks > 0x7fffffffll) return cudaErrorInvalidValue;
    k_topology_check<<<(unsigned)(tiles * chunks), TOPO_NT, 0, st>>>(X, A, N, rcov, extra, expect, a, b, c, dmax * dmax,
                                                                      dmax, tiles, out, events, cap, n0, n1);
    return cudaGetLastError();
}

// ddrax.cart_to_dipolar_from_df (ddrax.py:110-126) as a stand-alone call: pair vectors -> the 11 real words
// r^-3, Y20, Re/Im Y21, Re/Im Y22, F20, Re/Im F21, Re/Im F22 (rows of out[11][n] in that order)
__global__ void k_cart_to_dipolar(const double* __restrict__ dx, const double* __restrict__ dy,
                                  const double* __restrict__ dz, const double* __restrict__ dr, i64 n,
                                  double* __restrict__ out) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (i64)gridDim.x * blockDim.x) {
        const double r = dr[k];
        const DDWords w = dd_words(dx[k], dy[k], dz[k], r * r);
        out[k] = w.r3;
        out[n + k] = w.y20;
        out[2 * n + k] = w.y21r; out[3 * n + k] = w.y21i;
        out[4 * n + k] = w.y22r; out[5 * n + k] = w.y22i;
        out[6 * n + k] = w.f20;
        out[7 * n + k] = w.f21r; out[8 * n + k] = w.f21i;
        out[9 * n + k] = w.f22r; out[10 * n + k] = w.f22i;
    }
}
cudaError_t launch_cart_to_dipolar(const double* dx, const double* dy, const double* dz, const double* dr, i64 n,
                                   double* out, int n_sm, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    i64 blocks = (n + 255) / 256;
    if (blocks > 16 * n_sm) blocks = 16 * n_sm;
    k_cart_to_dipolar<<<(unsigned)blocks, 256, 0, st>>>(dx, dy, dz, dr, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ ragged dipolar series
// One CTA per pair: compress the frames in which the pair is present (r2 < rmax^2) into 7
// complex series of length N_p (Y20, Y21, Y22, r^-3, F20, F21, F22; real ones with zero
// imaginary plane) and record rank -> frame.  W[pair][7][2][N], fidx[pair][N], lens[7*pair+s].
__global__ void __launch_bounds__(256)
k_dd_ragged_words(const double* __restrict__ X, i64 N, i64 NC, const int* __restrict__ pi, const int* __restrict__ pj,
                  i64 pair0, double a, double b, double c, double rmax2, double* __restrict__ W,
                  int* __restrict__ fidx, i64* __restrict__ lens, double* __restrict__ scale,
                  double* __restrict__ bias, const unsigned char* __restrict__ mask) {
    // E sub-rows of 256 frames per round and ONE barrier per round (the per-warp counts are double-buffered and
    // every thread carries the running output position itself): the kernel is bound by load latency and
    // barriers, not by bandwidth
    constexpr int E = 4;
    __shared__ int wcnt[2][E][8];
    __shared__ double sh[8];
    const i64 pl = blockIdx.x;
    const i64 p = pair0 + pl;
    const double* xi = X + (i64)pi[p] * 3 * N;
    const double* xj = X + (i64)pj[p] * 3 * N;
    double* w = W + pl * 14 * NC;   // NC = capacity of a compressed series (<= N): pairs of one length class
    int* fi = fidx + pl * NC;
    double rsum = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    i64 base = 0;
    int buf = 0;
    for (i64 n0 = 0; n0 < N; n0 += 256 * E, buf ^= 1) {
        double dx[E], dy[E], dz[E], r2[E];
        unsigned ball[E];
        bool ok[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const i64 n = n0 + e * 256 + threadIdx.x;
            ok[e] = false;
            dx[e] = dy[e] = dz[e] = 0.0;
            r2[e] = 1.0;
            if (n < N) {
                double sx, sy, sz;
                axis_min(__ldg(xi + n), __ldg(xj + n), a, dx[e], sx);
                axis_min(__ldg(xi + N + n), __ldg(xj + N + n), b, dy[e], sy);
                axis_min(__ldg(xi + 2 * N + n), __ldg(xj + 2 * N + n), c, dz[e], sz);
                r2[e] = __dadd_rn(__dadd_rn(sx, sy), sz);
                ok[e] = r2[e] < rmax2 && (!mask || mask[p * N + n]);  // mask: frames in which the pair row passes the molecule filter
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ball[e] = __ballot_sync(0xffffffffu, ok[e]);
            if (lane == 0) wcnt[buf][e][warp] = __popc(ball[e]);
        }
        __syncthreads();
        i64 off = base;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int mine = 0, tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int cq = wcnt[buf][e][q];
                tot += cq;
                if (q < warp) mine += cq;
            }
            const i64 o = off + mine + __popc(ball[e] & below);
            if (ok[e] && o < NC) {
                const DDWords d = dd_words(dx[e], dy[e], dz[e], r2[e]);
                fi[o] = (int)(n0 + e * 256 + threadIdx.x);
                w[0 * NC + o] = d.y20;  w[1 * NC + o] = 0.0;
                w[2 * NC + o] = d.y21r; w[3 * NC + o] = d.y21i;
                w[4 * NC + o] = d.y22r; w[5 * NC + o] = d.y22i;
                w[6 * NC + o] = d.r3;   w[7 * NC + o] = 0.0;
                w[8 * NC + o] = d.f20;  w[9 * NC + o] = 0.0;
                w[10 * NC + o] = d.f21r; w[11 * NC + o] = d.f21i;
                w[12 * NC + o] = d.f22r; w[13 * NC + o] = d.f22i;
                rsum += d.r3;
            }
            off += tot;
        }
        base = off;
    }
    const double t = block_sum_256(rsum, sh);
    if (threadIdx.x == 0) {
        const i64 np = base < NC ? base : NC;   // a pair longer than the class capacity is a caller error: truncated
        for (int s = 0; s < 7; ++s) {
            lens[pl * 7 + s] = np;
            scale[pl * 7 + s] = 1.0;
        }
        bias[2 * pl] = t;      // sum of r^-3 over present frames (slot 3, re plane)
        bias[2 * pl + 1] = 0.0;
    }
}
cudaError_t launch_dd_ragged_words(const double* X, i64 N, i64 NC, const int* pi, const int* pj, i64 pair0, i64 npairs,
                                   double a, double b, double c, double rmax, double* W, int* fidx, i64* lens,
                                   double* scale, double* bias, const unsigned char* mask, cudaStream_t st) {
    if (npairs <= 0) return cudaSuccess;
    k_dd_ragged_words<<<(unsigned)npairs, 256, 0, st>>>(X, N, NC, pi, pj, pair0, a, b, c, rmax * rmax, W, fidx, lens,
                                                        scale, bias, mask);
    return cudaGetLastError();
}

// per pair of a ragged batch: number of frames it is present in, and the frames hit by any pair
__global__ void k_dd_ragged_presence(const int* __restrict__ fidx, const i64* __restrict__ lens, i64 N,
                                     i64* __restrict__ pair_frames, int* __restrict__ frame_hit) {
    const i64 pl = blockIdx.x;
    const i64 np = lens[pl * 7];
    if (threadIdx.x == 0 && pair_frames) pair_frames[pl] = np;
    if (frame_hit)
        for (i64 k = threadIdx.x; k < np; k += blockDim.x) frame_hit[fidx[pl * N + k]] = 1;
}
cudaError_t launch_dd_ragged_presence(const int* fidx, const i64* lens, i64 N, i64 npairs, i64* pair_frames,
                                      int* frame_hit, cudaStream_t st) {
    if (npairs <= 0 || (!pair_frames && !frame_hit)) return cudaSuccess;
    k_dd_ragged_presence<<<(unsigned)npairs, 256, 0, st>>>(fidx, lens, N, pair_frames, frame_hit);
    return cudaGetLastError();
}

// Packing of the even power spectra of the ragged pairs, two per inverse transform (rows 2t, 2t + 1 of the packed
// array share transform t).  Two spectra in one transform leak into each other at the level of the rounding error of
// the LARGER one, and the r^-3-weighted series (columns 3..6, ~ r^-6) are many orders of magnitude below the plain
// harmonics (columns 0..2, O(1)): like is packed with like.  Per block of two pairs (series o = 0..13, o = 7 pair + c):
//   (0,1) (2,9) (3,6) (4,5) (7,8) (10,13) (11,12)
// A last odd pair takes 8 rows; row 3 of its block (series 9 does not exist) is zero.
__device__ __constant__ signed char c_pos2orig[14] = {0, 1, 2, 9, 3, 6, 4, 5, 7, 8, 10, 13, 11, 12};
__device__ __constant__ signed char c_orig2pos[14] = {0, 1, 2, 4, 6, 7, 5, 8, 9, 3, 10, 12, 13, 11};
__device__ __forceinline__ i64 dd_pack_row_of_series(i64 s) { return (s / 14) * 14 + c_orig2pos[s % 14]; }
__device__ __forceinline__ i64 dd_pack_series_of_row(i64 r) { return (r / 14) * 14 + c_pos2orig[r % 14]; }

// G[acc][fidx[pair][k]] += acf[row][k] / N_p : scatter of per-pair ACFs to frame bins.
// acf holds Re(FFT(even P))[k] * (1/M) for k < NC, one row per packed row (series = pair*7 + acc).
__global__ void k_dd_ragged_scatter(const double* __restrict__ acf, const int* __restrict__ fidx,
                                    const i64* __restrict__ lens, i64 N, i64 NC, i64 nfft, double* __restrict__ G) {
    // rows on gridDim.x (limit 2^31 - 1; gridDim.y stops at 65535 = 9362 pairs), frame chunks on gridDim.y
    const i64 row = blockIdx.x;
    const i64 f = dd_pack_series_of_row(row);
    if (f >= nfft) return;
    const i64 pl = f / 7;
    const int acc = (int)(f % 7);
    const i64 np = lens[f];
    if (np <= 0) return;
    const double inv = 1.0 / (double)np;
    for (i64 k = (i64)blockIdx.y * blockDim.x + threadIdx.x; k < np; k += (i64)gridDim.y * blockDim.x)
        atomicAdd(G + (i64)acc * N + fidx[pl * NC + k], acf[row * NC + k] * inv);   // N frames per G row, NC per series
}
cudaError_t launch_dd_ragged_scatter(const double* acf, const int* fidx, const i64* lens, i64 N, i64 NC, i64 nfft,
                                     double* G, cudaStream_t st) {
    if (nfft <= 0) return cudaSuccess;
    const i64 rows = dd_pack_rows(nfft / 7);
    int gy = (int)((NC + 255) / 256);
    if (gy > 64) gy = 64;
    dim3 grid((unsigned)rows, gy);
    k_dd_ragged_scatter<<<grid, 256, 0, st>>>(acf, fidx, lens, N, NC, nfft, G);
    return cudaGetLastError();
}

// Power spectra from the engine's internal order [f][k1][k2] (k = k1 + M1 k2) to natural order [f][k2][k1]:
// a tiled transpose, both sides coalesced.  The column pass of an inverse transform then reads its input
// with unit stride instead of gathering one 8-byte element per 32-byte sector.
__global__ void __launch_bounds__(256) k_spectrum_to_natural(const double* __restrict__ in, double* __restrict__ out,
                                                             int M1) {
    __shared__ double tile[32][33];
    const i64 f = blockIdx.z;
    in += f * (i64)M1 * DYNPY_ROW_LEN;
    out += f * (i64)M1 * DYNPY_ROW_LEN;
    const int k2_0 = blockIdx.x * 32, k1_0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int k1 = k1_0 + j;
        if (k1 < M1) tile[j][tx] = in[(i64)k1 * DYNPY_ROW_LEN + k2_0 + tx];
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int k1 = k1_0 + tx;
        if (k1 < M1) out[(i64)(k2_0 + j) * M1 + k1] = tile[tx][j];
    }
}
cudaError_t launch_spectrum_to_natural(const double* in, double* out, i64 n_spec, int M1, cudaStream_t st) {
    for (i64 f0 = 0; f0 < n_spec; f0 += 32768) {
        const i64 cnt = n_spec - f0 < 32768 ? n_spec - f0 : 32768;
        dim3 grid(DYNPY_ROW_LEN / 32, (unsigned)((M1 + 31) / 32), (unsigned)cnt);
        k_spectrum_to_natural<<<grid, 256, 0, st>>>(in + f0 * (i64)M1 * DYNPY_ROW_LEN, out + f0 * (i64)M1 * DYNPY_ROW_LEN, M1);
    }
    return cudaGetLastError();
}

// Even part of power spectra, natural order: out[f][k] = (P[f][k] + P[f][(M - k) % M]) / 2 with P in the engine's
// internal order (two-pass lengths: element k = k1 + M1 k2 at [k1][k2]; its mirror is [M1 - k1][4095 - k2], or
// [0][(4096 - k2) % 4096] for k1 = 0) or already natural (single-pass lengths).  Re FFT(P) = FFT(even part of P):
// the even parts of two spectra can share one transform (SpectrumPairLoader).
__global__ void __launch_bounds__(256) k_spectrum_even_natural(const double* __restrict__ in, double* __restrict__ out,
                                                               int M1, i64 f0) {
    __shared__ double tile[32][33];
    const i64 f = f0 + blockIdx.z;
    in += f * (i64)M1 * DYNPY_ROW_LEN;
    out += dd_pack_row_of_series(f) * (i64)M1 * DYNPY_ROW_LEN;
    const int k2_0 = blockIdx.x * 32, k1_0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int k1 = k1_0 + j, k2 = k2_0 + tx;
        if (k1 < M1) {
            const int m1 = k1 ? M1 - k1 : 0;
            const int m2 = k1 ? DYNPY_ROW_LEN - 1 - k2 : (DYNPY_ROW_LEN - k2) & (DYNPY_ROW_LEN - 1);
            tile[j][tx] = 0.5 * (in[(i64)k1 * DYNPY_ROW_LEN + k2] + in[(i64)m1 * DYNPY_ROW_LEN + m2]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int k1 = k1_0 + tx;
        if (k1 < M1) out[(i64)(k2_0 + j) * M1 + k1] = tile[tx][j];
    }
}
__global__ void k_spectrum_even_single(const double* __restrict__ in, double* __restrict__ out, i64 M, i64 total) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 f = e / M, k = e - f * M;
        out[dd_pack_row_of_series(f) * M + k] = 0.5 * (in[e] + in[f * M + (k ? M - k : 0)]);
    }
}
cudaError_t launch_spectrum_even_natural(const double* in, double* out, i64 n_spec, i64 M, cudaStream_t st) {
    if (n_spec <= 0) return cudaSuccess;
    if ((n_spec / 7) & 1) {   // last odd pair: the partner row of its column 2 does not exist
        cudaError_t e = cudaMemsetAsync(out + (14 * (n_spec / 14) + 3) * M, 0, (size_t)M * sizeof(double), st);
        if (e != cudaSuccess) return e;
    }
    if (M <= DYNPY_ROW_LEN) {
        const i64 total = n_spec * M;
        i64 blocks = (total + 255) / 256;
        if (blocks > 65535 * 16) blocks = 65535 * 16;
        k_spectrum_even_single<<<(unsigned)blocks, 256, 0, st>>>(in, out, M, total);
        return cudaGetLastError();
    }
    const int M1 = (int)(M / DYNPY_ROW_LEN);
    for (i64 f0 = 0; f0 < n_spec; f0 += 32768) {
        const i64 cnt = n_spec - f0 < 32768 ? n_spec - f0 : 32768;
        dim3 grid(DYNPY_ROW_LEN / 32, (unsigned)((M1 + 31) / 32), (unsigned)cnt);
        k_spectrum_even_natural<<<grid, 256, 0, st>>>(in, out, M1, f0);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ spin-rotation
__device__ __forceinline__ void solve3(const double A_[3][3], const double b_[3], double x[3]) {
    // Gaussian elimination with partial pivoting (what numpy.linalg.solve / LAPACK gesv does)
    double A[3][3], b[3], ipiv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { b[i] = b_[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = A_[i][j]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double best = fabs(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < 3; ++i) if (fabs(A[i][k]) > best) { best = fabs(A[i][k]); piv = i; }
#pragma unroll
        for (int i = k + 1; i < 3; ++i) if (piv == i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) { double t = A[k][j]; A[k][j] = A[i][j]; A[i][j] = t; }
            double t = b[k]; b[k] = b[i]; b[i] = t;
        }
        ipiv[k] = 1.0 / A[k][k];
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            const double l = A[i][k] * ipiv[k];
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j >= k) A[i][j] -= l * A[k][j];
            b[i] -= l * b[k];
        }
    }
    // back substitution with the reciprocals of the pivots already at hand (three divisions fewer; one rounding of
    // difference to b / A, far inside the tolerance: this kernel is bound by fp64 issue)
    x[2] = b[2] * ipiv[2];
    x[1] = (b[1] - A[1][2] * x[2]) * ipiv[1];
    x[0] = (b[0] - A[0][1] * x[1] - A[0][2] * x[2]) * ipiv[0];
}
__device__ __forceinline__ void inv3(const double m[3][3], double o[3][3]) {
    const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    const double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
    const double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    const double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
    const double id = 1.0 / det;
    o[0][0] = c00 * id; o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
    o[1][0] = c01 * id; o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
    o[2][0] = c02 * id; o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
}
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void normalize3(double v[3]) {
    // v / |v| through rsqrt + two Newton steps (error ~1 ulp, like sqrt + 3 divisions at a fifth of the instructions)
    const double s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double r = rsqrt(s);
    r = r * fma(-0.5 * s, r * r, 1.5);
    r = r * fma(-0.5 * s, r * r, 1.5);
    v[0] *= r; v[1] *= r; v[2] *= r;
}

// thread = (frame, molecule); SR_func1 (SRrax.py:438-510):
//   r, v relative to the mass-weighted mean; R = sum(r^2 1 - r r^T) unweighted; I mass-weighted;
//   L' = sum r x v unweighted; omega = solve(R, L'); body axes from two minimum-image pair
//   vectors; omega_b = ax^-1 omega; I_b = ax^-1 I ax; J = I_b omega_b — evaluated as omega_b = ax^T omega,
//   J = ax^T (I omega), which is the same to rounding because ax is orthonormal.
// NAT > 0: atoms per molecule known at compile time (positions and velocities stay in registers between
// the centre-of-mass pass and the tensor pass); NAT == 0: run-time count, two passes over memory.
template <int NAT>
__global__ void __launch_bounds__(128)
k_sr_angmom(const double* __restrict__ pos, const double* __restrict__ vel, i64 N, const int* __restrict__ mol_atoms,
            i64 n_mol, int nat_rt, const double* __restrict__ mass, const int* __restrict__ axis_atoms, int axis_rule,
            double ca, double cb, double cc, double ica, double icb, double icc, double inv_dt,
            i64 n_out, double* __restrict__ J, double* __restrict__ omega, double* __restrict__ axes) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 m = blockIdx.y;
    const int nat = NAT > 0 ? NAT : nat_rt;
    // the per-slot masses are the same for every molecule: 1 / sum formed by one thread per CTA (one division
    // instead of 128; the additions run in the order of the per-thread sum)
    __shared__ double s_inv_msum;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < nat; ++k) s += __ldg(mass + k);
        s_inv_msum = 1.0 / s;
    }
    __syncthreads();
    if (t >= n_out) return;
    const i64 fr = vel ? t : t + 1;  // frame whose positions are used
    const int* at = mol_atoms + m * nat;
    constexpr int NR = NAT > 0 ? NAT : 1;
    double rx[NR][3], vx[NR][3], mk_[NR];
    auto fetch = [&](int k, double (&x)[3], double (&v)[3]) {
        const double* p = pos + (i64)__ldg(at + k) * 3 * N + fr;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            x[d] = p[d * N];
            v[d] = vel ? vel[((i64)__ldg(at + k) * 3 + d) * N + fr] : (x[d] - p[d * N - 1]) * inv_dt;
        }
    };
    double msum = 0.0, cp[3] = {0, 0, 0}, cv[3] = {0, 0, 0};
    if constexpr (NAT > 0) {
#pragma unroll
        for (int k = 0; k < NAT; ++k) {
            fetch(k, rx[k], vx[k]);
            mk_[k] = __ldg(mass + k);
        }
#pragma unroll
        for (int k = 0; k < NAT; ++k) {
            msum += mk_[k];
#pragma unroll
            for (int d = 0; d < 3; ++d) { cp[d] += mk_[k] * rx[k][d]; cv[d] += mk_[k] * vx[k][d]; }
        }
    } else {
        for (int k = 0; k < nat; ++k) {
            double x[3], v[3];
            fetch(k, x, v);
            const double mk = mass[k];
            msum += mk;
#pragma unroll
            for (int d = 0; d < 3; ++d) { cp[d] += mk * x[d]; cv[d] += mk * v[d]; }
        }
    }
    const double im = s_inv_msum;
#pragma unroll
    for (int d = 0; d < 3; ++d) { cp[d] *= im; cv[d] *= im; }
    double R[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, I[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rv[3] = {0, 0, 0};
    auto tensor_terms = [&](const double (&x)[3], const double (&vv)[3], double mk) {
        double r[3], v[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { r[d] = x[d] - cp[d]; v[d] = vv[d] - cv[d]; }
        const double xx = r[0] * r[0], yy = r[1] * r[1], zz = r[2] * r[2];
        R[0][0] += yy + zz; R[1][1] += xx + zz; R[2][2] += xx + yy;
        R[0][1] += -r[0] * r[1]; R[0][2] += -r[0] * r[2]; R[1][2] += -r[1] * r[2];
        I[0][0] += mk * (yy + zz); I[1][1] += mk * (xx + zz); I[2][2] += mk * (xx + yy);
        I[0][1] += -mk * r[0] * r[1]; I[0][2] += -mk * r[0] * r[2]; I[1][2] += -mk * r[1] * r[2];
        double c3[3];
        cross3(r, v, c3);
        rv[0] += c3[0]; rv[1] += c3[1]; rv[2] += c3[2];
    };
    if constexpr (NAT > 0) {
#pragma unroll
        for (int k = 0; k < NAT; ++k) tensor_terms(rx[k], vx[k], mk_[k]);
    } else {
        for (int k = 0; k < nat; ++k) {
            double x[3], v[3];
            fetch(k, x, v);
            tensor_terms(x, v, mass[k]);
        }
    }
    R[1][0] = R[0][1]; R[2][0] = R[0][2]; R[2][1] = R[1][2];
    I[1][0] = I[0][1]; I[2][0] = I[0][2]; I[2][1] = I[1][2];
    double o[3];
    solve3(R, rv, o);
    // body axes from the two pair vectors.  The reference takes them from the pair table of the WRAPPED
    // coordinates (d = u_a + s L - u_b, the image s with the smallest |d|); for two bonded atoms that is the
    // nearest image of the raw difference, d = (x_a - x_b) - L rint((x_a - x_b)/L), to one rounding of
    // |x| * 2^-53 — far inside the tolerance, and 12 exact fmods + 18 candidate evaluations cheaper (this
    // kernel is bound by fp64 issue, not by HBM).
    const int* ax4 = axis_atoms + m * 4;
    const double cell[3] = {ca, cb, cc};
    const double icell[3] = {ica, icb, icc};
    double dvec[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const double* pa = pos + (i64)ax4[2 * h] * 3 * N + fr;
        const double* pb = pos + (i64)ax4[2 * h + 1] * 3 * N + fr;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double dd = pa[d * N] - pb[d * N];
            dvec[h][d] = fma(-cell[d], rint(dd * icell[d]), dd);
        }
    }
    double X[3], Y[3], Z[3];
    if (axis_rule == 1) {
        Z[0] = -(dvec[0][0] + dvec[1][0]); Z[1] = -(dvec[0][1] + dvec[1][1]); Z[2] = -(dvec[0][2] + dvec[1][2]);
        normalize3(Z);
        cross3(Z, dvec[0], X); normalize3(X);
        cross3(Z, X, Y); normalize3(Y);
    } else {
        Z[0] = dvec[0][0]; Z[1] = dvec[0][1]; Z[2] = dvec[0][2];
        normalize3(Z);
        cross3(Z, dvec[1], X); normalize3(X);
        cross3(Z, X, Y); normalize3(Y);
    }
    // ax = [x y z] as columns is orthonormal to rounding (x, y are normalised cross products with z), so
    // ax^-1 = ax^T and  J = I_b omega_b = ax^T I ax ax^T omega = ax^T (I omega):  18 products instead of the
    // inverse and two 3x3 matrix products the reference spells out (la.inv, np.matmul; SRrax.py:491-498)
    const double ax[3][3] = {{X[0], Y[0], Z[0]}, {X[1], Y[1], Z[1]}, {X[2], Y[2], Z[2]}};  // columns x,y,z
    double ob[3], Io[3], Jb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Io[i] = I[i][0] * o[0] + I[i][1] * o[1] + I[i][2] * o[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ob[i] = ax[0][i] * o[0] + ax[1][i] * o[1] + ax[2][i] * o[2];
        Jb[i] = ax[0][i] * Io[0] + ax[1][i] * Io[1] + ax[2][i] * Io[2];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (J) J[(m * 3 + d) * n_out + t] = Jb[d];
        if (omega) omega[(m * 3 + d) * n_out + t] = ob[d];
    }
    if (axes) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) axes[(m * 9 + i * 3 + j) * n_out + t] = ax[i][j];
    }
}
cudaError_t launch_sr_angmom(const double* pos, const double* vel, i64 N, const int* mol_atoms, i64 n_mol, int nat,
                             const double* mass, const int* axis_atoms, int axis_rule, double a, double b, double c,
                             double inv_dt, double* J, double* omega, double* axes, cudaStream_t st) {
    const i64 n_out = vel ? N : N - 1;
    if (n_out <= 0 || n_mol <= 0) return cudaSuccess;
    for (i64 m0 = 0; m0 < n_mol; m0 += 65535) {
        i64 cnt = n_mol - m0 < 65535 ? n_mol - m0 : 65535;
        dim3 grid((unsigned)((n_out + 127) / 128), (unsigned)cnt);
#define SR_LAUNCH(NATC)                                                                                             \
    k_sr_angmom<NATC><<<grid, 128, 0, st>>>(pos, vel, N, mol_atoms + m0 * nat, cnt, nat, mass, axis_atoms + m0 * 4,   \
                                            axis_rule, a, b, c, 1.0 / a, 1.0 / b, 1.0 / c, inv_dt, n_out,                   \
                                            J ? J + m0 * 3 * n_out : nullptr,                                           \
                                            omega ? omega + m0 * 3 * n_out : nullptr,                                   \
                                            axes ? axes + m0 * 9 * n_out : nullptr)
        switch (nat) {  // water, methyl / ammonia-like, methane, acetonitrile; anything else: run-time loop
            case 3: SR_LAUNCH(3); break;
            case 4: SR_LAUNCH(4); break;
            case 5: SR_LAUNCH(5); break;
            case 6: SR_LAUNCH(6); break;
            default: SR_LAUNCH(0); break;
        }
#undef SR_LAUNCH
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ Simpson
// scipy.integrate.simpson(y, x=x): composite Simpson with the non-uniform-h weights of
// scipy's _basic_simpson on an odd number of points + the even-N end rule (mode).
__device__ double simpson_odd_block(const double* __restrict__ y, const double* __restrict__ x, i64 start, i64 npts,
                                    double* sh) {
    double acc = 0.0;
    const i64 nseg = (npts - 1) / 2;
    for (i64 s = threadIdx.x; s < nseg; s += blockDim.x) {
        const i64 i0 = start + 2 * s;
        const double h0 = x[i0 + 1] - x[i0], h1 = x[i0 + 2] - x[i0 + 1];
        const double hsum = h0 + h1, hprod = h0 * h1;
        const double h0divh1 = (h1 != 0.0) ? h0 / h1 : 0.0;
        const double t0 = (h0divh1 != 0.0) ? 1.0 / h0divh1 : 0.0;
        const double t1 = (hprod != 0.0) ? hsum / hprod : 0.0;
        acc += hsum / 6.0 * (y[i0] * (2.0 - t0) + y[i0 + 1] * (hsum * t1) + y[i0 + 2] * (2.0 - h0divh1));
    }
    // block reduction (blockDim = 1024)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}
__global__ void __launch_bounds__(1024)
k_simpson(const double* __restrict__ Y, const double* __restrict__ x, i64 n, i64 col_stride, int mode,
          double* __restrict__ out) {
    __shared__ double sh[32];
    const double* y = Y + (i64)blockIdx.x * col_stride;
    double res = 0.0;
    if (n == 1) {
        res = 0.0;
    } else if (n & 1) {
        res = simpson_odd_block(y, x, 0, n, sh);
    } else if (n == 2) {
        res = 0.5 * (x[1] - x[0]) * (y[1] + y[0]);
    } else if (mode == 1) {
        const double A = simpson_odd_block(y, x, 0, n - 1, sh);
        const double B = simpson_odd_block(y, x, 1, n - 1, sh);
        const double ta = 0.5 * (x[n - 1] - x[n - 2]) * (y[n - 1] + y[n - 2]);
        const double tb = 0.5 * (x[1] - x[0]) * (y[1] + y[0]);
        res = 0.5 * ((A + ta) + (B + tb));
    } else {
        res = simpson_odd_block(y, x, 0, n - 1, sh);
        const double h0 = x[n - 2] - x[n - 3], h1 = x[n - 1] - x[n - 2];
        double den = 6.0 * (h1 + h0);
        const double alpha = den != 0.0 ? (2.0 * h1 * h1 + 3.0 * h0 * h1) / den : 0.0;
        den = 6.0 * h0;
        const double beta = den != 0.0 ? (h1 * h1 + 3.0 * h0 * h1) / den : 0.0;
        den = 6.0 * h0 * (h0 + h1);
        const double eta = den != 0.0 ? (h1 * h1 * h1) / den : 0.0;
        res = res + alpha * y[n - 1] + beta * y[n - 2] - eta * y[n - 3];
    }
    if (threadIdx.x == 0) out[blockIdx.x] = res;
}
cudaError_t launch_simpson(const double* y, const double* x, i64 n, int n_cols, i64 col_stride, int mode,
                           double* out, cudaStream_t st) {
    if (n_cols <= 0) return cudaSuccess;
    k_simpson<<<n_cols, 1024, 0, st>>>(y, x, n, col_stride, mode, out);
    return cudaGetLastError();
}

}  // namespace dynpy
