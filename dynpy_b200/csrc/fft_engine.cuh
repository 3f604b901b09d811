// dynpy_b200 — batched fp64 forward-FFT engine for the Wiener–Khinchin hot path (sm_100a).
//
// Replaces the per-series  np.fft.fft(f, n=2N) ; |X|^2 ; np.fft.ifft  of
// wiener_khinchin (reference qrax.py:147-156 == ddrax.py:276-285 == SRrax.py:216-225):
// here every series gets ONE forward transform whose |X|^2 is accumulated over the
// series of a batch (linearity of the inverse transform), and one transform per output
// column turns the accumulated spectrum into the summed ACF at the very end.
//
// Decomposition (checked against numpy in tools/fft_model.py):
//   M <= 4096 : one kernel (k_rows) does the whole transform in registers + shared memory.
//   M  > 4096 : four-step split M = M1 x 4096.  k_cols does the M1-point column FFTs
//               (input stride 4096, zero padding folded into the loader), multiplies by the
//               inter-pass twiddle W_M^{n2 k1} and writes T[k1][n2]; k_rows does the
//               4096-point row FFTs and consumes X[k1 + M1 k2] straight from registers.
//   Inside a kernel: in-place multi-step DIF, radix <= 16 per step in registers
//               (bit-reversed register outputs are free: compile-time indexing), steps
//               exchange through padded shared memory, per-step twiddles W_P^{k m} are
//               powers of one table entry.
// The accumulated spectrum lives in "internal order"  S'[k1*M2 + k2]  <->  k = k1 + M1*k2;
// only the final transform (loader kind SPECTRUM) un-permutes it.
//
// Bandwidth/FLOP accounting per kernel is in DESIGN.md §kernels.
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>

#include "codelets.cuh"

namespace dynpy {

typedef double2 cd;
typedef long long i64;

#define DYNPY_ROW_LEN 4096  // M2 of the two-pass split

// transform lengths the engine implements: powers of two in [32, 2^24] and M1 * 4096 for the mixed-radix
// column plans M1 = 20, 24, 40, 50 (thread-per-column kernels with 10-, 12-, 20- and 25-point codelets)
static inline bool fft_length_ok(long long M) {
    if (M >= 32 && M <= (1ll << 24) && (M & (M - 1)) == 0) return true;
    return M == 20ll * DYNPY_ROW_LEN || M == 24ll * DYNPY_ROW_LEN || M == 40ll * DYNPY_ROW_LEN || M == 50ll * DYNPY_ROW_LEN;
}
// smallest implemented length >= 2n - 1 (any such M gives the reference's linear ACF, qrax.py:153)
static inline long long fft_length_for(long long n) {
    long long m = 32;
    while (m < 2 * n) m <<= 1;
    const int mixed[4] = {20, 24, 40, 50};
    for (int k = 0; k < 4; ++k) {
        const long long c = (long long)mixed[k] * DYNPY_ROW_LEN;
        if (c >= 2 * n - 1 && c < m) m = c;
    }
    return m;
}
#ifndef DYNPY_ROWS_MINB
#define DYNPY_ROWS_MINB 2   // resident CTAs per SM the row kernel is compiled for
#endif
#ifndef DYNPY_COLS_MINB
#define DYNPY_COLS_MINB 2   // same for the column kernel
#endif

__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cd csq(cd a) {
    return make_double2(fma(a.x, a.x, -(a.y * a.y)), (a.x + a.x) * a.y);
}

// ---------------------------------------------------------------- constant twiddles W_32^t
__device__ __forceinline__ double c32(int t) {  // cos(2 pi t / 32), t in [0,8]
    switch (t) {
        case 0: return 1.0;
        case 1: return 0.98078528040323044913;
        case 2: return 0.92387953251128675613;
        case 3: return 0.83146961230254523708;
        case 4: return 0.70710678118654752440;
        case 5: return 0.55557023301960222474;
        case 6: return 0.38268343236508977173;
        case 7: return 0.19509032201612826785;
        default: return 0.0;
    }
}
// d * W_32^t, t compile-time after unrolling (0 <= t < 16)
__device__ __forceinline__ cd mul_w32(cd d, int t) {
    if (t == 0) return d;
    if (t == 8) return make_double2(d.y, -d.x);
    if (t == 4) {
        const double r = 0.70710678118654752440;
        return make_double2((d.x + d.y) * r, (d.y - d.x) * r);
    }
    if (t == 12) {
        const double r = 0.70710678118654752440;
        return make_double2((d.y - d.x) * r, -(d.x + d.y) * r);
    }
    // W = c - i s with c = cos(2 pi t/32), s = sin(2 pi t/32)
    double c, s;
    if (t < 8) { c = c32(t); s = c32(8 - t); }
    else       { c = -c32(16 - t); s = c32(t - 8); }
    return make_double2(fma(d.x, c, d.y * s), fma(d.y, c, -(d.x * s)));
}

__host__ __device__ constexpr int brev(int k, int R) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (k & 1); k >>= 1; }
    return r;
}

// radix-R DIF in registers: natural order in, X[k] ends in v[brev(k,R)]
template <int R>
__device__ __forceinline__ void fft_dif(cd (&v)[R]) {
    constexpr int LG = (R >= 16) ? 4 : (R >= 8) ? 3 : (R >= 4) ? 2 : (R >= 2) ? 1 : 0;
#pragma unroll
    for (int st = 0; st < LG; ++st) {
        const int len = R >> st;
        const int half = len >> 1;
#pragma unroll
        for (int g = 0; g < R; g += len) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                cd a = v[g + j], b = v[g + j + half];
                v[g + j] = cadd(a, b);
                v[g + j + half] = mul_w32(csub(a, b), j * (32 / len));
            }
        }
    }
}
template <>
__device__ __forceinline__ void fft_dif<1>(cd (&v)[1]) {}

// v[brev(k)] *= w^k  for k = 1..R-1 (binary powering keeps the error depth logarithmic)
template <int R>
__device__ __forceinline__ void apply_powers(cd (&v)[R], cd w) {
    if (R < 2) return;
    cd p[R];
    p[0] = make_double2(1.0, 0.0);
    p[1] = w;
#pragma unroll
    for (int k = 2; k < R; ++k) p[k] = (k & 1) ? cmul(p[k - 1], w) : csq(p[k >> 1]);
#pragma unroll
    for (int k = 1; k < R; ++k) v[brev(k, R)] = cmul(v[brev(k, R)], p[k]);
}

// ---------------------------------------------------------------- plans
template <int L> struct Plan;
#define DYNPY_PLAN(LEN, A, B, C, NS) \
    template <> struct Plan<LEN> { static constexpr int R1 = A, R2 = B, R3 = C, S = NS, RL = (NS == 1 ? A : (NS == 2 ? B : C)); }
DYNPY_PLAN(1, 1, 1, 1, 1);
DYNPY_PLAN(2, 2, 1, 1, 1);
DYNPY_PLAN(4, 4, 1, 1, 1);
DYNPY_PLAN(8, 8, 1, 1, 1);
DYNPY_PLAN(16, 16, 1, 1, 1);
DYNPY_PLAN(32, 2, 16, 1, 2);
DYNPY_PLAN(64, 4, 16, 1, 2);
DYNPY_PLAN(128, 8, 16, 1, 2);
DYNPY_PLAN(256, 16, 16, 1, 2);
DYNPY_PLAN(512, 2, 16, 16, 3);
DYNPY_PLAN(1024, 4, 16, 16, 3);
DYNPY_PLAN(2048, 8, 16, 16, 3);
DYNPY_PLAN(4096, 16, 16, 16, 3);

__host__ __device__ constexpr int pad16(int p) { return p + (p >> 4); }

// twiddle table W_4096^j: a per-device global array filled once by k_init_twiddles and passed
// to every kernel as `twp`
__device__ __forceinline__ cd ldg_cd(const cd* p) {
    double2 r;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
#define tw4096(j) ldg_cd(twp + (j))

// W_M^e for 0 <= e < M (M power of two <= 2^40): one accurate sincospi, used only for
// loop-invariant twiddles computed once per thread
__device__ __forceinline__ cd w_exact(i64 e, double inv_M) {
    double s, c;
    sincospi(2.0 * (double)e * inv_M, &s, &c);
    return make_double2(c, -s);
}

// ---------------------------------------------------------------- loaders
enum LoaderKind { LOAD_PLANAR = 0, LOAD_QR = 1, LOAD_SPECTRUM = 2 };

// Planar real/imag series.  FFT f reads  re[(f/nslots)*hi + (f%nslots)*lo + n]  and the same
// offset of `im` (nullptr: purely real; FFTs with f >= n_im_valid also read a zero imaginary
// part — the odd series of a real-pair packing).  bias[2*item], bias[2*item+1] (optional) hold
// the sums over n of the re / im series of slot `bias_slot`; value - sum/N is transformed
// (G_r = ACF of r^-3 - mean(r^-3), ddrax.py:292).
struct PlanarLoader {
    static constexpr bool DENSE = false;  // zero-padded series: samples n >= N are structurally zero
    const double* re;
    const double* im;
    i64 hi, lo, n_im_valid;
    int nslots;
    i64 N;
    const i64* lens;      // optional per-FFT length (ragged pairs); nullptr -> N
    const double* scale;  // optional per-FFT scale; nullptr -> 1
    const double* bias;   // optional [2 * n_items] sums of the re / im series of slot `bias_slot`
    int bias_slot;
    struct Ctx { const double* r; const double* i; i64 n; double sc, br, bi; };
    __device__ __forceinline__ Ctx begin(i64 f) const {
        Ctx c;
        i64 base = (f / nslots) * hi + (f % nslots) * lo;
        c.r = re + base;
        c.i = (im != nullptr && f < n_im_valid) ? im + base : nullptr;
        c.n = lens ? lens[f] : N;
        c.sc = scale ? scale[f] : 1.0;
        c.br = 0.0; c.bi = 0.0;
        if (bias && (int)(f % nslots) == bias_slot) {
            c.br = bias[2 * (f / nslots)] / (double)c.n;
            c.bi = bias[2 * (f / nslots) + 1] / (double)c.n;
        }
        return c;
    }
    __device__ __forceinline__ cd load(const Ctx& c, i64 n) const {
        if (n >= c.n) return make_double2(0.0, 0.0);
        double a = __ldg(c.r + n) - c.br;
        double b = c.i ? (__ldg(c.i + n) - c.bi) : 0.0;
        return make_double2(a * c.sc, b * c.sc);
    }
    // two-phase form: raw() only issues the loads (so a whole batch can be in flight), finish()
    // does the arithmetic.  Out-of-range samples read as the bias, i.e. 0 after finish().
    __device__ __forceinline__ cd raw(const Ctx& c, i64 n) const {
        cd v = make_double2(c.br, c.bi);
        if (n < c.n) {
            v.x = __ldg(c.r + n);
            if (c.i) v.y = __ldg(c.i + n);
        }
        return v;
    }
    __device__ __forceinline__ cd finish(const Ctx& c, cd v) const {
        return make_double2((v.x - c.br) * c.sc, (v.y - c.bi) * c.sc);
    }
};

// Quadrupolar loader (qrax.py:115-145): EFG stored [nucleus][6][N] with component order
// Vxx,Vyy,Vzz,Vxy,Vxz,Vyz.  Five FFT slots per nucleus PAIR (a,b):
//   0: R20(a) + i R20(b)           (real series of two nuclei share one transform)
//   1: R2,-1(a) = Vxz - i Vyz      2: R2,-2(a) = (Vxx-Vyy)/2 - i Vxy
//   3: R2,-1(b)                    4: R2,-2(b)
// |R2,+m|^2-spectra equal those of R2,-m (conjugate / sign), so +m columns reuse them.
struct QrLoader {
    static constexpr bool DENSE = false;
    const double* V;
    i64 N;       // frames
    i64 nnuc;    // nuclei in this shard
    struct Ctx { const double* a; const double* b; int slot; };
    __device__ __forceinline__ Ctx begin(i64 f) const {
        Ctx c;
        i64 pr = f / 5;
        c.slot = (int)(f % 5);
        i64 na = 2 * pr, nb = 2 * pr + 1;
        c.a = V + na * 6 * N;
        c.b = (nb < nnuc) ? V + nb * 6 * N : nullptr;
        return c;
    }
    __device__ __forceinline__ cd load(const Ctx& c, i64 n) const {
        if (n >= N) return make_double2(0.0, 0.0);
        const double k20 = 1.2247448713915889;  // 3*sqrt(1/6)
        if (c.slot == 0) {
            double a = k20 * __ldg(c.a + 2 * N + n);
            double b = c.b ? k20 * __ldg(c.b + 2 * N + n) : 0.0;
            return make_double2(a, b);
        }
        const double* p = (c.slot <= 2) ? c.a : c.b;
        if (!p) return make_double2(0.0, 0.0);
        if (c.slot == 1 || c.slot == 3) return make_double2(__ldg(p + 4 * N + n), -__ldg(p + 5 * N + n));
        return make_double2(0.5 * (__ldg(p + n) - __ldg(p + N + n)), -__ldg(p + 3 * N + n));
    }
    __device__ __forceinline__ cd raw(const Ctx& c, i64 n) const { return load(c, n); }
    __device__ __forceinline__ cd finish(const Ctx& c, cd v) const { return v; }
    // The same sample with the kind of series fixed at compile time (the slot is uniform over a CTA: the column
    // kernel branches once per series instead of once per sample).  K = 0: k20 (Vzz_a + i Vzz_b); K = 1: R2,+-1 =
    // Vxz - i Vyz; K = 2: R2,+-2 = (Vxx - Vyy) / 2 - i Vxy, of nucleus p (nullptr: the missing second nucleus).
    template <int K>
    __device__ __forceinline__ cd load_kind(const Ctx& c, const double* p, i64 n) const {
        cd v = make_double2(0.0, 0.0);
        if (n < N) {
            if constexpr (K == 0) {
                const double k20 = 1.2247448713915889;
                v.x = k20 * __ldg(c.a + 2 * N + n);
                if (c.b) v.y = k20 * __ldg(c.b + 2 * N + n);
            } else if constexpr (K == 1) {
                if (p) v = make_double2(__ldg(p + 4 * N + n), -__ldg(p + 5 * N + n));
            } else {
                if (p) v = make_double2(0.5 * (__ldg(p + n) - __ldg(p + N + n)), -__ldg(p + 3 * N + n));
            }
        }
        return v;
    }
};

// Final transform: reads the accumulated spectrum S' (internal order) of accumulator f as a
// purely real length-M series in natural frequency order.
struct SpectrumLoader {
    static constexpr bool DENSE = true;   // a full-length spectrum
    const double* S;
    i64 M;
    int M1;  // 1 for single-pass spectra
    int natural;  // 1: S is in natural frequency order (after k_spectrum_to_natural); 0: the engine's internal order
    struct Ctx { const double* s; };
    __device__ __forceinline__ Ctx begin(i64 f) const { Ctx c; c.s = S + f * M; return c; }
    __device__ __forceinline__ cd load(const Ctx& c, i64 n) const {
        // internal order: element n sits at [n % M1][n / M1] — consecutive columns of the inverse's column pass hit
        // different 32-byte sectors (a gather); spectra that are transformed back one by one are transposed first
        i64 idx = (M1 > 1 && !natural) ? (n % M1) * (i64)DYNPY_ROW_LEN + n / M1 : n;
        return make_double2(__ldg(c.s + idx), 0.0);
    }
    __device__ __forceinline__ cd raw(const Ctx& c, i64 n) const { return load(c, n); }
    __device__ __forceinline__ cd finish(const Ctx& c, cd v) const { return v; }
};

// Two real EVEN spectra per transform, natural frequency order (k_spectrum_to_natural with sym = 1): series
// P[2f] + i P[2f+1].  The transform of a real even sequence is real, so Re / Im of the result are the two ACFs
// (EPI_WRITE_REAL2): half the inverse transforms of the per-series path.
struct SpectrumPairLoader {
    static constexpr bool DENSE = true;
    const double* S;   // [n_spec][M]
    i64 M;
    i64 n_spec;
    struct Ctx { const double* a; const double* b; };
    __device__ __forceinline__ Ctx begin(i64 f) const {
        Ctx c; c.a = S + 2 * f * M; c.b = (2 * f + 1 < n_spec) ? S + (2 * f + 1) * M : nullptr; return c;
    }
    __device__ __forceinline__ cd load(const Ctx& c, i64 n) const {
        return make_double2(__ldg(c.a + n), c.b ? __ldg(c.b + n) : 0.0);
    }
    __device__ __forceinline__ cd raw(const Ctx& c, i64 n) const { return load(c, n); }
    __device__ __forceinline__ cd finish(const Ctx& c, cd v) const { return v; }
};

// Row loader for the second pass: T[f_local][k1][n2]
struct TLoader {
    const cd* T;
    int M1;
    i64 f_base;
    struct Ctx { const cd* row; };
    __device__ __forceinline__ Ctx begin(i64 f, int k1) const {
        Ctx c; c.row = T + ((f - f_base) * M1 + k1) * (i64)DYNPY_ROW_LEN; return c;
    }
    __device__ __forceinline__ cd load(const Ctx& c, i64 n) const { return c.row[n]; }
};

// ---------------------------------------------------------------- epilogue parameters
enum EpiMode { EPI_ACC_POWER = 0, EPI_WRITE_REAL = 1, EPI_WRITE_POWER = 2, EPI_WRITE_REAL2 = 3 };
struct EpiParams {
    double* out;        // ACC_POWER: S'[acc][M] ; WRITE_REAL: out[f][n_out] ; WRITE_POWER: P[f][M] (internal order)
    i64 M;              // full transform length
    int nslots;         // ACC_POWER: acc id = 4-bit field (f % nslots) of acc_pack
    unsigned long long acc_pack;
    i64 n_out;          // WRITE_REAL: outputs kept (k < n_out)
    i64 n_series;       // WRITE_REAL2: out[2f] = Re, out[2f+1] = Im (when 2f+1 < n_series)
    double scale;       // WRITE_REAL: multiplier
};

// ---------------------------------------------------------------- second pass / single pass
// One CTA = one row of length L (two-pass: row k1 = blockIdx.x of every FFT in
// [f0 + blockIdx.y*stride ...]; single pass: the whole transform, M1 == 1).
// Threads: L / RL (one last-step butterfly each), padded up to a warp.
template <int L, bool TWO_PASS, class Loader, int MODE>
__global__ void __launch_bounds__((L / Plan<L>::RL) < 32 ? 32 : (L / Plan<L>::RL), DYNPY_ROWS_MINB)
k_rows(Loader ld, EpiParams ep, const cd* __restrict__ twp, i64 f_begin, i64 f_end, int f_step, int M1) {
    typedef Plan<L> P;
    constexpr int R1 = P::R1, R2 = P::R2, R3 = P::R3, S = P::S, RL = P::RL;
    constexpr int NT = L / RL;
    constexpr int P1 = L / R1, P2 = P1 / R2;
    extern __shared__ double2 smem[];
    const int tid = threadIdx.x;
    const bool active = tid < NT;
    const int k1row = TWO_PASS ? blockIdx.x : 0;
    // f index handled by this CTA: f = f_begin + slot + nslots*(item) pattern is produced by the host as
    // f = f_begin + blockIdx.y + j*f_step
    double acc[RL];
#pragma unroll
    for (int j = 0; j < RL; ++j) acc[j] = 0.0;

    for (i64 f = f_begin + blockIdx.y; f < f_end; f += f_step) {
        cd v[16];
        // ---------------- step 1 : inputs from the loader
        typename Loader::Ctx ctx;
        if constexpr (TWO_PASS) ctx = ld.begin(f, k1row); else ctx = ld.begin(f);
        if (active) {
#pragma unroll 1
            for (int it = 0; it < RL / R1; ++it) {
                const int m = tid + it * NT;
                cd u[R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) u[q] = ld.load(ctx, (i64)(q * P1 + m));
                fft_dif<R1>(u);
                if constexpr (S == 1) {
#pragma unroll
                    for (int k = 0; k < R1; ++k) v[k] = u[brev(k, R1)];
                } else {
                    apply_powers<R1>(u, tw4096(m * (4096 / L)));
#pragma unroll
                    for (int k = 0; k < R1; ++k) smem[pad16(k * P1 + m)] = u[brev(k, R1)];
                }
            }
        }
        int kk0 = 0;    // natural index of v[0] inside the row; v[j] <-> kk0 + j*kstride
        int kstride = 1;
        if constexpr (S >= 2) {
            __syncthreads();
            if (active) {
#pragma unroll 1
                for (int it = 0; it < RL / R2; ++it) {
                    const int b = tid + it * NT;
                    const int h = b / P2, m = b % P2;
                    cd u[R2];
#pragma unroll
                    for (int q = 0; q < R2; ++q) u[q] = smem[pad16(h * P1 + q * P2 + m)];
                    fft_dif<R2>(u);
                    if constexpr (S == 2) {
#pragma unroll
                        for (int k = 0; k < R2; ++k) v[k] = u[brev(k, R2)];
                        kk0 = h; kstride = R1;
                    } else {
                        apply_powers<R2>(u, tw4096(m * (4096 / P1)));
#pragma unroll
                        for (int k = 0; k < R2; ++k) smem[pad16(h * P1 + k * P2 + m)] = u[brev(k, R2)];
                    }
                }
            }
        }
        if constexpr (S == 3) {
            __syncthreads();
            if (active) {
                const int h = tid;
                cd u[R3];
#pragma unroll
                for (int q = 0; q < R3; ++q) u[q] = smem[pad16(h * R3 + q)];
                fft_dif<R3>(u);
#pragma unroll
                for (int k = 0; k < R3; ++k) v[k] = u[brev(k, R3)];
                kk0 = (h / R2) + R1 * (h % R2); kstride = R1 * R2;
            }
        }
        // ---------------- epilogue: v[j] = X_row[kk0 + j*kstride], j < RL
        if constexpr (MODE == EPI_ACC_POWER) {
#pragma unroll
            for (int j = 0; j < RL; ++j) acc[j] = fma(v[j].x, v[j].x, fma(v[j].y, v[j].y, acc[j]));
        } else if constexpr (MODE == EPI_WRITE_REAL) {
            if (active) {
#pragma unroll
                for (int j = 0; j < RL; ++j) {
                    i64 k = (i64)k1row + (i64)M1 * (kk0 + j * kstride);
                    if (k < ep.n_out) ep.out[f * ep.n_out + k] = v[j].x * ep.scale;
                }
            }
        } else if constexpr (MODE == EPI_WRITE_REAL2) {
            if (active) {
                const bool two = 2 * f + 1 < ep.n_series;
#pragma unroll
                for (int j = 0; j < RL; ++j) {
                    i64 k = (i64)k1row + (i64)M1 * (kk0 + j * kstride);
                    if (k < ep.n_out) {
                        ep.out[2 * f * ep.n_out + k] = v[j].x * ep.scale;
                        if (two) ep.out[(2 * f + 1) * ep.n_out + k] = v[j].y * ep.scale;
                    }
                }
            }
        } else {  // EPI_WRITE_POWER: per-series |X|^2 in internal order
            if (active) {
#pragma unroll
                for (int j = 0; j < RL; ++j)
                    ep.out[f * ep.M + (i64)k1row * L + kk0 + j * kstride] = fma(v[j].x, v[j].x, v[j].y * v[j].y);
            }
        }
        if constexpr (S >= 2) __syncthreads();  // smem is rewritten by the next transform's step 1
        // carry the mapping out of the loop for the flush (identical every iteration)
        if constexpr (MODE == EPI_ACC_POWER) {
            if (f + f_step >= f_end) {
                // flush through shared memory so the global reduction is coalesced
                double* sr = reinterpret_cast<double*>(smem);
                if (active) {
#pragma unroll
                    for (int j = 0; j < RL; ++j) sr[kk0 + j * kstride] = acc[j];
                }
                __syncthreads();
                const int slot = (int)((f_begin + blockIdx.y) % ep.nslots);
                double* dst = ep.out + (i64)((ep.acc_pack >> (4 * slot)) & 15ull) * ep.M + (i64)k1row * L;
                for (int k = tid; k < L; k += blockDim.x) atomicAdd(dst + k, sr[k]);
            }
        }
    }
}

// ---------------------------------------------------------------- first pass (columns)
// CTA = tile of C consecutive columns n2 of every FFT f = f_begin + blockIdx.y + j*f_step.
// Threads: (M1/RL) * C, column index fastest (coalesced 8*C-byte loads, 16*C-byte stores).
template <int M1, int C, class Loader>
__global__ void __launch_bounds__((M1 / Plan<M1>::RL) * C, DYNPY_COLS_MINB)
k_cols(Loader ld, cd* __restrict__ T, const cd* __restrict__ twp, i64 M, double inv_M, i64 f_begin, i64 f_end,
       int f_step) {
    typedef Plan<M1> P;
    constexpr int R1 = P::R1, R2 = P::R2, R3 = P::R3, S = P::S, RL = P::RL;
    constexpr int NB = M1 / RL;  // last-step butterflies per column
    constexpr int P1 = M1 / R1, P2 = P1 / R2;
    constexpr int M2 = DYNPY_ROW_LEN;
    extern __shared__ double2 smem[];
    const int c = threadIdx.x % C;
    const int bb = threadIdx.x / C;
    const int n2 = blockIdx.x * C + c;

    // loop-invariant inter-pass twiddles for this thread's outputs k1_j = kk0 + j*kstride
    int kk0, kstride;
    if constexpr (S == 1) { kk0 = 0; kstride = 1; }
    else if constexpr (S == 2) { kk0 = bb; kstride = R1; }
    else { kk0 = (bb / R2) + R1 * (bb % R2); kstride = R1 * R2; }
    cd tw[RL];
    {
        const cd base = w_exact(((i64)n2 * kk0) % M, inv_M);
        const cd stp = w_exact(((i64)n2 * kstride) % M, inv_M);
        tw[0] = base;
        if (RL > 1) tw[1] = stp;
#pragma unroll
        for (int j = 2; j < RL; ++j) tw[j] = (j & 1) ? cmul(tw[j - 1], stp) : csq(tw[j >> 1]);
#pragma unroll
        for (int j = 1; j < RL; ++j) tw[j] = cmul(tw[j], base);
    }

    for (i64 f = f_begin + blockIdx.y; f < f_end; f += f_step) {
        typename Loader::Ctx ctx = ld.begin(f);
        cd v[RL];
        cd in[RL];  // all RL input samples of this thread are requested before any is consumed
#pragma unroll
        for (int it = 0; it < RL / R1; ++it) {
#pragma unroll
            for (int q = 0; q < R1; ++q)
                in[it * R1 + q] = ld.raw(ctx, (i64)(q * P1 + bb + it * NB) * M2 + n2);
        }
#pragma unroll
        for (int it = 0; it < RL / R1; ++it) {
            const int m = bb + it * NB;
            cd u[R1];
#pragma unroll
            for (int q = 0; q < R1; ++q) u[q] = ld.finish(ctx, in[it * R1 + q]);
            fft_dif<R1>(u);
            if constexpr (S == 1) {
#pragma unroll
                for (int k = 0; k < R1; ++k) v[k] = u[brev(k, R1)];
            } else {
                apply_powers<R1>(u, tw4096(m * (4096 / M1)));
#pragma unroll
                for (int k = 0; k < R1; ++k) smem[pad16(k * P1 + m) * C + c] = u[brev(k, R1)];
            }
        }
        if constexpr (S >= 2) {
            __syncthreads();
#pragma unroll 1
            for (int it = 0; it < RL / R2; ++it) {
                const int b = bb + it * NB;
                const int h = b / P2, m = b % P2;
                cd u[R2];
#pragma unroll
                for (int q = 0; q < R2; ++q) u[q] = smem[pad16(h * P1 + q * P2 + m) * C + c];
                fft_dif<R2>(u);
                if constexpr (S == 2) {
#pragma unroll
                    for (int k = 0; k < R2; ++k) v[k] = u[brev(k, R2)];
                } else {
                    apply_powers<R2>(u, tw4096(m * (4096 / P1)));
#pragma unroll
                    for (int k = 0; k < R2; ++k) smem[pad16(h * P1 + k * P2 + m) * C + c] = u[brev(k, R2)];
                }
            }
        }
        if constexpr (S == 3) {
            __syncthreads();
            const int h = bb;
            cd u[R3];
#pragma unroll
            for (int q = 0; q < R3; ++q) u[q] = smem[pad16(h * R3 + q) * C + c];
            fft_dif<R3>(u);
#pragma unroll
            for (int k = 0; k < R3; ++k) v[k] = u[brev(k, R3)];
        }
        cd* dst = T + ((f - f_begin) * M1) * (i64)M2 + n2;
#pragma unroll
        for (int j = 0; j < RL; ++j) dst[(i64)(kk0 + j * kstride) * M2] = cmul(v[j], tw[j]);
        if constexpr (S >= 2) __syncthreads();
    }
}

// ---------------------------------------------------------------- first pass, one thread per column
// Any loader, M1 in {16, 20, 24, 32, 40, 50}: the M1-point column transform is a radix-2 DIF split into two
// H = M1/2 point codelets done in registers (even k1: x[j] + x[j+H]; odd k1: (x[j] - x[j+H]) W_M1^j), one
// after the other to stay inside the register file; the inter-pass twiddles W_M^{n2 k1} run as a chain
// per parity.  Zero-padded loaders (series of N <= M/2 samples: Loader::DENSE == false) never touch
// the upper half.  No shared memory, no barrier; consecutive threads are consecutive columns.
// M = 204800 = 50 x 4096 is the transform length for 65536 < N <= 102400 frames.
template <int H> struct HalfColumn;
template <> struct HalfColumn<8> {
    static __device__ __forceinline__ void even(cd (&v)[8]) { fft8(v); }
    static __device__ __forceinline__ void odd(cd (&v)[8]) { fft8_odd(v); }
};
template <> struct HalfColumn<16> {
    static __device__ __forceinline__ void even(cd (&v)[16]) { fft16(v); }
    static __device__ __forceinline__ void odd(cd (&v)[16]) { fft16_odd(v); }
};
template <> struct HalfColumn<10> {
    static __device__ __forceinline__ void even(cd (&v)[10]) { fft10(v); }
    static __device__ __forceinline__ void odd(cd (&v)[10]) { fft10_odd(v); }
};
template <> struct HalfColumn<12> {
    static __device__ __forceinline__ void even(cd (&v)[12]) { fft12(v); }
    static __device__ __forceinline__ void odd(cd (&v)[12]) { fft12_odd(v); }
};
template <> struct HalfColumn<20> {
    static __device__ __forceinline__ void even(cd (&v)[20]) { fft20(v); }
    static __device__ __forceinline__ void odd(cd (&v)[20]) { fft20_odd(v); }
};
template <> struct HalfColumn<25> {
    static __device__ __forceinline__ void even(cd (&v)[25]) { fft25(v); }
    static __device__ __forceinline__ void odd(cd (&v)[25]) { fft25_odd(v); }
};

template <int M1, class Loader>
__global__ void __launch_bounds__(256, 2) k_cols_tpcg(Loader ld, cd* __restrict__ T, double inv_M, i64 f_begin,
                                                      i64 f_end, int f_step) {
    constexpr int M2 = DYNPY_ROW_LEN;
    constexpr int H = M1 / 2;
    extern __shared__ double2 smem[];
    cd* stash = smem + threadIdx.x;  // zero-padded loaders: the H samples of this thread, kept for the odd half
    const int n2 = blockIdx.x * blockDim.x + threadIdx.x;
    const cd w1 = w_exact(n2, inv_M);          // W_M^{n2}
    const cd w2 = w_exact(2 * (i64)n2, inv_M);  // W_M^{2 n2}
    for (i64 f = f_begin + blockIdx.y; f < f_end; f += f_step) {
        typename Loader::Ctx ctx = ld.begin(f);
        cd* dst = T + ((f - f_begin) * M1) * (i64)M2 + n2;
#pragma unroll 1
        for (int par = 0; par < 2; ++par) {
            cd v[H];
            if constexpr (std::is_same<Loader, QrLoader>::value) {
                if (par == 0) {
                    const double* pn = (ctx.slot <= 2) ? ctx.a : ctx.b;
                    auto fill = [&](auto kind) {
                        constexpr int K = decltype(kind)::value;
#pragma unroll
                        for (int j = 0; j < H; ++j) v[j] = ld.template load_kind<K>(ctx, pn, (i64)j * M2 + n2);
                    };
                    if (ctx.slot == 0) fill(std::integral_constant<int, 0>{});
                    else if (ctx.slot & 1) fill(std::integral_constant<int, 1>{});
                    else fill(std::integral_constant<int, 2>{});
#pragma unroll
                    for (int j = 0; j < H; ++j) stash[j * 256] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < H; ++j) v[j] = stash[j * 256];
                }
            } else if (Loader::DENSE || par == 0) {
#pragma unroll
                for (int j = 0; j < H; ++j) v[j] = ld.raw(ctx, (i64)j * M2 + n2);
#pragma unroll
                for (int j = 0; j < H; ++j) v[j] = ld.finish(ctx, v[j]);
                if constexpr (Loader::DENSE) {
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        const cd b = ld.finish(ctx, ld.raw(ctx, (i64)(j + H) * M2 + n2));
                        v[j] = par ? csub(v[j], b) : cadd(v[j], b);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < H; ++j) stash[j * 256] = v[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < H; ++j) v[j] = stash[j * 256];
            }
            cd w;
            if (par) { HalfColumn<H>::odd(v); w = w1; }
            else { HalfColumn<H>::even(v); w = make_double2(1.0, 0.0); }
#pragma unroll
            for (int a = 0; a < H; ++a) {
                dst[(i64)(2 * a + par) * M2] = cmul(v[a], w);
                if (a + 1 < H) w = cmul(w, w2);
            }
        }
    }
}

}  // namespace dynpy
