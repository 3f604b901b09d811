// dynpy_b200 — finite-frequency spectral densities, sm_100a.
//
//   dft_run            numpy.fft.fft of arbitrary length n (ddrax.fourier_T, ddrax.py:169-173: the ACF
//                      columns have the trajectory's length, not a power of two) by Bluestein's chirp-z
//                      identity  X_k = w_k * sum_j (x_j w_j) conj(w_{k-j}),  w_j = exp(-i pi j^2 / n),
//                      i.e. one circular convolution of length M = 2^m >= 2n - 1, done with an
//                      out-of-place Stockham radix-2 transform (autosort: no bit reversal).  Chirp and
//                      butterfly factors come from sincospi of exactly reduced arguments (j^2 mod 2n in
//                      integers), so the error stays at the few-ulp level of an ordinary FFT.
//   cutoff_bounds      integration bound of spec_dens_extr_narr / spectral_dens with cutoff=True
//                      (ddrax.py:188-198, qrax.py:172-182): first index whose trailing `window`-mean of
//                      y / y[0] is within `tol` of zero.
//
// A handful of series of the trajectory's length per run: bandwidth-trivial next to the correlation path,
// so the kernels are plain grid-stride code.
#include "common.h"
#include "kernels.h"

namespace dynpy {

__device__ __forceinline__ cd blu_chirp(i64 j, i64 n) {  // exp(-i pi j^2 / n), j < 2^31
    const i64 r = (j * j) % (2 * n);
    double s, c;
    sincospi(-(double)r / (double)n, &s, &c);
    return make_double2(c, s);
}

// a[j] = x[j] * w_j for j < n, zero up to M
__global__ void k_blu_load(const double* __restrict__ re, const double* __restrict__ im, i64 n, i64 M, cd* __restrict__ a) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (i64)gridDim.x * blockDim.x) {
        cd v = make_double2(0.0, 0.0);
        if (j < n) v = cmul(make_double2(re[j], im ? im[j] : 0.0), blu_chirp(j, n));
        a[j] = v;
    }
}

// b[j] = conj(w_j) at j and M - j (the chirp is even in j), zero elsewhere
__global__ void k_blu_chirp(i64 n, i64 M, cd* __restrict__ b) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (i64)gridDim.x * blockDim.x) {
        const i64 d = j < n ? j : (M - j < n ? M - j : -1);
        cd v = make_double2(0.0, 0.0);
        if (d >= 0) { v = blu_chirp(d, n); v.y = -v.y; }
        b[j] = v;
    }
}

// one Stockham radix-2 pass: Ns = 1, 2, 4, ..., M/2
__global__ void k_stockham2(const cd* __restrict__ in, cd* __restrict__ out, i64 M, i64 Ns) {
    const i64 half = M >> 1;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += (i64)gridDim.x * blockDim.x) {
        const i64 k = j & (Ns - 1);
        double s, c;
        sincospi(-(double)k / (double)Ns, &s, &c);
        const cd v0 = in[j];
        const cd v1 = cmul(in[j + half], make_double2(c, s));
        const i64 j0 = ((j - k) << 1) + k;
        out[j0] = make_double2(v0.x + v1.x, v0.y + v1.y);
        out[j0 + Ns] = make_double2(v0.x - v1.x, v0.y - v1.y);
    }
}

// c = conj(A * B): the inverse transform is run as conj(FFT(conj(.)))
__global__ void k_blu_mul(const cd* __restrict__ A, const cd* __restrict__ B, i64 M, cd* __restrict__ c) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (i64)gridDim.x * blockDim.x) {
        const cd v = cmul(A[j], B[j]);
        c[j] = make_double2(v.x, -v.y);
    }
}

// X_k = w_k * conj(c_k) / M
__global__ void k_blu_out(const cd* __restrict__ c, i64 n, i64 n_out, double inv_M, cd* __restrict__ out) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n_out; k += (i64)gridDim.x * blockDim.x) {
        const cd v = c[k];
        const cd r = cmul(make_double2(v.x * inv_M, -v.y * inv_M), blu_chirp(k, n));
        out[k] = r;
    }
}

static i64 blu_length(i64 n) {
    i64 M = 2;
    while (M < 2 * n - 1) M <<= 1;
    return M;
}

i64 dft_workspace_bytes(i64 n) { return 3 * blu_length(n) * (i64)sizeof(cd) + 256; }

// in-place-looking power-of-two transform of `a` using `tmp`; returns the buffer holding the result
static cudaError_t stockham(cd*& a, cd*& tmp, i64 M, int grid, cudaStream_t st) {
    for (i64 Ns = 1; Ns < M; Ns <<= 1) {
        k_stockham2<<<grid, 256, 0, st>>>(a, tmp, M, Ns);
        cd* t = a; a = tmp; tmp = t;
    }
    return cudaGetLastError();
}

int dft_run(const double* re, const double* im, i64 n_series, i64 n, cd* out, i64 n_out, void* ws, i64 ws_bytes,
            int n_sm, cudaStream_t st) {
    const i64 M = blu_length(n);
    if (ws_bytes < dft_workspace_bytes(n)) {
        set_error("dynpy_dft_f64: workspace too small (need %lld bytes)", (long long)dft_workspace_bytes(n));
        return DYNPY_ERR_WORKSPACE;
    }
    cd* B = (cd*)ws;
    cd* a = B + M;
    cd* tmp = a + M;
    const int grid = 8 * n_sm;
    // chirp spectrum (once per call); after an even number of passes the result may sit in either buffer
    k_blu_chirp<<<grid, 256, 0, st>>>(n, M, a);
    DYNPY_CHECK(stockham(a, tmp, M, grid, st), "dft chirp transform");
    { cd* t = B; B = a; a = t; }  // keep the chirp spectrum where it landed, recycle the spare buffer
    for (i64 s = 0; s < n_series; ++s) {
        k_blu_load<<<grid, 256, 0, st>>>(re + s * n, im ? im + s * n : nullptr, n, M, a);
        DYNPY_CHECK(stockham(a, tmp, M, grid, st), "dft forward transform");
        k_blu_mul<<<grid, 256, 0, st>>>(a, B, M, a);
        DYNPY_CHECK(stockham(a, tmp, M, grid, st), "dft inverse transform");
        k_blu_out<<<grid, 256, 0, st>>>(a, n, n_out, 1.0 / (double)M, out + s * n_out);
        DYNPY_CHECK(cudaGetLastError(), "dft output");
    }
    return DYNPY_OK;
}

// bounds[col] = first i in [window-1, n-1) with |mean(y[i-window+1 .. i] / y[0])| <= tol, else n-1
// (np.isclose(rolling(window).mean(), 0, atol=tol): NaN for incomplete windows never matches).
__global__ void k_cutoff_init(unsigned long long* bounds, int n_cols, i64 n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cols) bounds[c] = (unsigned long long)(n - 1);
}
__global__ void k_cutoff_bound(const double* __restrict__ Y, i64 n, i64 col_stride, int window, double tol,
                               unsigned long long* __restrict__ bounds) {
    const double* y = Y + (i64)blockIdx.y * col_stride;
    const double y0 = y[0];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x + (window - 1); i < n - 1; i += (i64)gridDim.x * blockDim.x) {
        if ((unsigned long long)i >= bounds[blockIdx.y]) break;  // a smaller bound is already known (monotone hint only)
        double s = 0.0;
        for (int k = window - 1; k >= 0; --k) s += y[i - k] / y0;
        const double m = s / (double)window;
        if (fabs(m) <= tol) atomicMin(bounds + blockIdx.y, (unsigned long long)i);
    }
}

cudaError_t launch_cutoff_bounds(const double* y, i64 n, int n_cols, i64 col_stride, int window, double tol,
                                 i64* bounds, int n_sm, cudaStream_t st) {
    if (n_cols <= 0) return cudaSuccess;
    k_cutoff_init<<<(n_cols + 63) / 64, 64, 0, st>>>((unsigned long long*)bounds, n_cols, n);
    i64 blocks = (n + 255) / 256;
    if (blocks > 8 * n_sm) blocks = 8 * n_sm;
    if (blocks < 1) blocks = 1;
    k_cutoff_bound<<<dim3((unsigned)blocks, (unsigned)n_cols), 256, 0, st>>>(y, n, col_stride, window, tol,
                                                                             (unsigned long long*)bounds);
    return cudaGetLastError();
}

}  // namespace dynpy
