// dynpy_b200 — persistent producer/consumer kernel for the fused dipolar path (sm_100a).
//
// Same mathematics as k_dd_words + k_cols + k_rows (kernels_misc.cu, fft_engine.cuh), organised so
// that neither the words Z nor the four-step intermediate T ever travel to HBM:
//   * producer CTAs own a fixed tile of C columns n2 (inter-pass twiddles live in registers),
//     generate the pair words for their samples on the fly from the wrapped coordinates,
//     run the M1-point column FFTs and write their tile of T into a ring slot;
//   * consumer CTAs own a fixed row k1, wait until every tile of a transform has landed, run the
//     4096-point row FFT out of L2 and accumulate |X|^2 in registers;
//   * the ring holds W transforms (W*M*16 bytes, sized to stay inside L2); slots are handed over
//     with monotonic counters (release: __threadfence + atomicAdd, acquire: ld.acquire.gpu);
//   * transforms are streamed in accumulator-major order (all Y20 transforms of the call, then all
//     Y21 ...), so a consumer flushes its registers into S' only 7 times per call.
// All CTAs must be co-resident (cooperative launch); every spin has a bounded budget and raises
// an error flag instead of hanging.
#pragma once
#include "dd_device.cuh"
#include "fft_engine.cuh"

namespace dynpy {

template <int M1> struct Plan8;  // column plans whose last radix is 8
template <> struct Plan8<16> { static constexpr int R1 = 2; };
template <> struct Plan8<32> { static constexpr int R1 = 4; };
template <> struct Plan8<64> { static constexpr int R1 = 8; };

struct StreamParams {
    const double* X;     // wrapped coords [A][3][N]
    const int* pi;
    const int* pj;
    i64 n_pairs;         // P
    i64 N;               // frames
    double a, b, c;      // cell
    double* S;           // [7][M] accumulators (internal order)
    i64 M;
    double inv_M;
    cd* T;               // ring: W slots of M points
    int W;
    unsigned long long* ready;  // [W]
    unsigned long long* done;   // [W]
    unsigned long long* total;  // [1] rows of type-0 transforms consumed so far
    double* rsum;        // [P] sums of r^-3 (filled while streaming type 0, read by type 3)
    int* err;
    const cd* twp;       // W_4096^j
    i64 bnd[8];          // sequence boundaries of the 7 accumulator types
    int n_prod, kp, kc;  // producer CTAs, producer copies per tile, consumer copies per row
    long long spin_limit;
    long long* dbg;      // optional [gridDim.x][4]: cycles waiting (a), waiting (b), total, iterations
};

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ cd ldcg_cd(const cd* p) {
    double2 r;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// thread 0 spins until *ctr >= target (bounded); everybody learns the outcome through smem
__device__ __forceinline__ bool wait_counter(const unsigned long long* ctr, unsigned long long target,
                                             long long limit, int* err, int* sflag, long long* waited = nullptr) {
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        long long n = 0;
        int ok = 1;
        while (ld_acquire(ctr) < target) {
            __nanosleep(40);
            if (++n > limit || *((volatile int*)err) != 0) { ok = 0; atomicExch(err, 1); break; }
        }
        *sflag = ok;
    }
    __syncthreads();
    const bool ok = *sflag != 0;
    __syncthreads();
    if (waited) *waited += clock64() - t0;
    return ok;
}

// radix-R DIF whose upper half of inputs is structurally zero (zero padding: n1 >= M1/2)
template <int R>
__device__ __forceinline__ void fft_dif_halfzero(cd (&v)[R]) {
    if (R >= 2) {
        constexpr int half = R / 2;
#pragma unroll
        for (int j = 0; j < half; ++j) v[j + half] = mul_w32(v[j], j * (32 / R));
    }
    constexpr int LG = (R >= 16) ? 4 : (R >= 8) ? 3 : (R >= 4) ? 2 : (R >= 2) ? 1 : 0;
#pragma unroll
    for (int st = 1; st < LG; ++st) {
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

__device__ __forceinline__ int type_of_seq(const i64* bnd, i64 s) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < 7; ++k) t += (s >= bnd[k]) ? 1 : 0;
    return t;
}

// one pair-frame -> the complex sample of accumulator type `type` (packed types use .x only)
__device__ __forceinline__ void pair_sample(const StreamParams& p, i64 pair, i64 frame, double& dx, double& dy,
                                            double& dz, double& r2) {
    const double* xi = p.X + (i64)p.pi[pair] * 3 * p.N;
    const double* xj = p.X + (i64)p.pj[pair] * 3 * p.N;
    double sx, sy, sz;
    axis_min(__ldg(xi + frame), __ldg(xj + frame), p.a, dx, sx);
    axis_min(__ldg(xi + p.N + frame), __ldg(xj + p.N + frame), p.b, dy, sy);
    axis_min(__ldg(xi + 2 * p.N + frame), __ldg(xj + 2 * p.N + frame), p.c, dz, sz);
    r2 = __dadd_rn(__dadd_rn(sx, sy), sz);
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// (dx,dy,dz,r2) from six already loaded wrapped coordinates
__device__ __forceinline__ void pair_geom(const double (&v)[6], double a, double b, double c, double& dx, double& dy,
                                          double& dz, double& r2) {
    double sx, sy, sz;
    axis_min(v[0], v[3], a, dx, sx);
    axis_min(v[1], v[4], b, dy, sy);
    axis_min(v[2], v[5], c, dz, sz);
    r2 = __dadd_rn(__dadd_rn(sx, sy), sz);
}

template <int M1>
__global__ void __launch_bounds__(256, 2) k_dd_stream(const StreamParams p) {
    constexpr int R1 = Plan8<M1>::R1;
    constexpr int R2 = 8;
    constexpr int NB = M1 / 8;        // threads per column
    constexpr int C = 256 / NB;       // columns per producer tile
    constexpr int NPOS = 4096 / C;    // tile positions
    constexpr int P1 = 8;             // M1 / R1
    extern __shared__ double2 smem[];
    __shared__ int sflag;
    __shared__ double sred[16];
    const int tid = threadIdx.x;
    const i64 S_tot = p.bnd[7];
    const i64 slot_pts = p.M;
    long long w_a = 0, w_b = 0, n_it = 0;
    const long long t_start = clock64();

    if ((int)blockIdx.x < p.n_prod) {
        // ------------------------------------------------------------------ producer
        const int pos = blockIdx.x % NPOS;
        const int copy = blockIdx.x / NPOS;
        const int c = tid % C;
        const int bb = tid / C;
        const int n2 = pos * C + c;
        cd tw[R2];
        {
            const cd base = w_exact(((i64)n2 * bb) % p.M, p.inv_M);
            const cd stp = w_exact(((i64)n2 * R1) % p.M, p.inv_M);
            tw[0] = base;
            tw[1] = stp;
#pragma unroll
            for (int j = 2; j < R2; ++j) tw[j] = (j & 1) ? cmul(tw[j - 1], stp) : csq(tw[j >> 1]);
#pragma unroll
            for (int j = 1; j < R2; ++j) tw[j] = cmul(tw[j], base);
        }
        bool sums_ready = false;
        // thread-private staging area for the coordinates of the NEXT transform (cp.async, no
        // registers): value v of this thread lives at stage[v*256 + tid], v = frame_slot*6 + comp
        double* stage = reinterpret_cast<double*>(smem + pad16(M1) * C);
        constexpr int NFR = R2 / 2;               // frames per thread (upper half of n1 is padding)
        constexpr int QH = (R1 + 1) / 2;          // valid q per step-1 butterfly
        auto frame_of = [&](int fs) -> i64 {
            const int it = fs / QH, q = fs % QH;
            return (i64)(q * P1 + bb + it * NB) * 4096 + n2;
        };
        auto first_pair = [&](i64 sq) -> i64 {
            const int t = type_of_seq(p.bnd, sq);
            const i64 kk = sq - p.bnd[t];
            return (t == 0 || t == 3 || t == 4) ? 2 * kk : kk;
        };
        auto prefetch = [&](i64 sq) {
            if (sq < S_tot) {
                const i64 pr = first_pair(sq);
                const double* xi = p.X + (i64)p.pi[pr] * 3 * p.N;
                const double* xj = p.X + (i64)p.pj[pr] * 3 * p.N;
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) {
                    const i64 frame = frame_of(fs);
                    if (frame < p.N) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            cp_async8(stage + (fs * 6 + d) * 256 + tid, xi + d * p.N + frame);
                            cp_async8(stage + (fs * 6 + 3 + d) * 256 + tid, xj + d * p.N + frame);
                        }
                    }
                }
            }
            cp_async_commit();
        };
        prefetch(copy);
        for (i64 s = copy; s < S_tot; s += p.kp) {
            const int slot = (int)(s % p.W);
            const unsigned long long use = (unsigned long long)(s / p.W);
            const int type = type_of_seq(p.bnd, s);
            const i64 k = s - p.bnd[type];
            const bool packed = (type == 0 || type == 3 || type == 4);
            const i64 pa = packed ? 2 * k : k;
            const i64 pb = (packed && 2 * k + 1 < p.n_pairs) ? 2 * k + 1 : -1;
            // second pair of a packed transform: plain loads, issued before anything else
            double vb[NFR][6];
            if (pb >= 0) {
                const double* xi = p.X + (i64)p.pi[pb] * 3 * p.N;
                const double* xj = p.X + (i64)p.pj[pb] * 3 * p.N;
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) {
                    const i64 frame = frame_of(fs);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        vb[fs][d] = frame < p.N ? __ldg(xi + d * p.N + frame) : 0.0;
                        vb[fs][3 + d] = frame < p.N ? __ldg(xj + d * p.N + frame) : 1.0;
                    }
                }
            }
            double ma = 0.0, mb = 0.0;
            if (type == 3) {
                // the r^-3 sums were accumulated while type 0 streamed: make sure every type-0
                // transform has been produced AND consumed (happens-before through `total`)
                if (!sums_ready) {
                    if (!wait_counter(p.total, (unsigned long long)M1 * (unsigned long long)p.bnd[1], p.spin_limit,
                                      p.err, &sflag)) return;
                    sums_ready = true;
                }
                ma = __ldcg(p.rsum + pa) / (double)p.N;
                if (pb >= 0) mb = __ldcg(p.rsum + pb) / (double)p.N;
            }
            // staged coordinates of the first pair: out of shared memory into the samples
            cp_async_wait_all();
            cd smp[NFR];
            double racc_a = 0.0, racc_b = 0.0;
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) {
                smp[fs] = make_double2(0.0, 0.0);
                if (frame_of(fs) < p.N) {
                    double v[6];
#pragma unroll
                    for (int d = 0; d < 6; ++d) v[d] = stage[(fs * 6 + d) * 256 + tid];
                    double dx, dy, dz, r2;
                    pair_geom(v, p.a, p.b, p.c, dx, dy, dz, r2);
                    DDWords w = dd_words(dx, dy, dz, r2);
                    double re, im = 0.0;
                    switch (type) {
                        case 0: re = w.y20; break;
                        case 1: re = w.y21r; im = w.y21i; break;
                        case 2: re = w.y22r; im = w.y22i; break;
                        case 3: re = w.r3 - ma; break;
                        case 4: re = w.f20; break;
                        case 5: re = w.f21r; im = w.f21i; break;
                        default: re = w.f22r; im = w.f22i; break;
                    }
                    if (type == 0) racc_a += w.r3;
                    if (pb >= 0) {
                        pair_geom(vb[fs], p.a, p.b, p.c, dx, dy, dz, r2);
                        DDWords w2 = dd_words(dx, dy, dz, r2);
                        im = (type == 0) ? w2.y20 : (type == 3) ? (w2.r3 - mb) : w2.f20;
                        if (type == 0) racc_b += w2.r3;
                    }
                    smp[fs] = make_double2(re, im);
                }
            }
            prefetch(s + p.kp);  // the staging slots of this thread are free again
            if (use > 0) {
                if (!wait_counter(p.done + slot, (unsigned long long)M1 * use, p.spin_limit, p.err, &sflag, &w_a)) return;
            }
            ++n_it;
            // ---- step 1: radix-R1 (upper half of the inputs is zero padding)
#pragma unroll
            for (int it = 0; it < R2 / R1; ++it) {
                const int m = bb + it * NB;
                cd u[R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) u[q] = (q < QH) ? smp[it * QH + q] : make_double2(0.0, 0.0);
                if (R1 >= 2) fft_dif_halfzero<R1>(u);
                apply_powers<R1>(u, ldg_cd(p.twp + m * (4096 / M1)));
#pragma unroll
                for (int kq = 0; kq < R1; ++kq) smem[pad16(kq * P1 + m) * C + c] = u[brev(kq, R1)];
            }
            if (type == 0) {  // per-pair sums of r^-3 for the mean removal of type 3
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    racc_a += __shfl_down_sync(0xffffffffu, racc_a, o);
                    racc_b += __shfl_down_sync(0xffffffffu, racc_b, o);
                }
                if ((tid & 31) == 0) { sred[2 * (tid >> 5)] = racc_a; sred[2 * (tid >> 5) + 1] = racc_b; }
            }
            __syncthreads();
            if (type == 0 && tid == 0) {
                double ta = 0.0, tb = 0.0;
                for (int w8 = 0; w8 < 8; ++w8) { ta += sred[2 * w8]; tb += sred[2 * w8 + 1]; }
                atomicAdd(p.rsum + pa, ta);
                if (pb >= 0) atomicAdd(p.rsum + pb, tb);
            }
            // ---- step 2: radix-8 over the column, inter-pass twiddle, write the tile of T
            {
                cd u[R2];
#pragma unroll
                for (int q = 0; q < R2; ++q) u[q] = smem[pad16(bb * P1 + q) * C + c];
                fft_dif<R2>(u);
                cd* dst = p.T + (i64)slot * slot_pts + n2;
#pragma unroll
                for (int j = 0; j < R2; ++j) dst[(i64)(bb + R1 * j) * 4096] = cmul(u[brev(j, R2)], tw[j]);
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(p.ready + slot, 1ull);
            }
        }
    } else {
        // ------------------------------------------------------------------ consumer
        const int idx = blockIdx.x - p.n_prod;
        const int row = idx % M1;
        const int copy = idx / M1;
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0;
        int cur_type = -1;
        const int kk0 = (tid / 16) + 16 * (tid % 16);  // v[j] <-> row index kk0 + 256 j
        auto flush = [&](int type) {
            double* sr = reinterpret_cast<double*>(smem);
#pragma unroll
            for (int j = 0; j < 16; ++j) sr[kk0 + 256 * j] = acc[j];
            __syncthreads();
            double* dst = p.S + (i64)type * p.M + (i64)row * 4096;
            for (int k = tid; k < 4096; k += 256) atomicAdd(dst + k, sr[k]);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0;
        };
        for (i64 s = copy; s < S_tot; s += p.kc) {
            const int type = type_of_seq(p.bnd, s);
            if (type != cur_type) {
                if (cur_type >= 0) flush(cur_type);
                cur_type = type;
            }
            const int slot = (int)(s % p.W);
            const unsigned long long use = (unsigned long long)(s / p.W);
            if (!wait_counter(p.ready + slot, (unsigned long long)NPOS * (use + 1), p.spin_limit, p.err, &sflag, &w_a)) return;
            ++n_it;
            const cd* src = p.T + (i64)slot * slot_pts + (i64)row * 4096;
            cd u[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) u[q] = ldcg_cd(src + q * 256 + tid);
            fft_dif<16>(u);
            apply_powers<16>(u, ldg_cd(p.twp + tid));
#pragma unroll
            for (int k = 0; k < 16; ++k) smem[pad16(k * 256 + tid)] = u[brev(k, 16)];
            __syncthreads();
            if (tid == 0) {  // every thread of this CTA has its samples in registers: hand the row back
                __threadfence();
                atomicAdd(p.done + slot, 1ull);
                if (type == 0) atomicAdd(p.total, 1ull);
            }
            {
                const int h = tid / 16, m = tid % 16;
#pragma unroll
                for (int q = 0; q < 16; ++q) u[q] = smem[pad16(h * 256 + q * 16 + m)];
                fft_dif<16>(u);
                apply_powers<16>(u, ldg_cd(p.twp + m * 16));
#pragma unroll
                for (int k = 0; k < 16; ++k) smem[pad16(h * 256 + k * 16 + m)] = u[brev(k, 16)];
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; ++q) u[q] = smem[pad16(tid * 16 + q)];
            fft_dif<16>(u);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const cd x = u[brev(j, 16)];
                acc[j] = fma(x.x, x.x, fma(x.y, x.y, acc[j]));
            }
            __syncthreads();
        }
        if (cur_type >= 0) flush(cur_type);
    }
    if (p.dbg && tid == 0) {
        long long* d = p.dbg + 4 * (long long)blockIdx.x;
        d[0] = w_a; d[1] = w_b; d[2] = clock64() - t_start; d[3] = n_it;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused word generator + column pass for the batched two-kernel path.  A CTA owns a tile of C
// columns n2 (inter-pass twiddles in registers) and loops over the pair-pairs ("items") of the
// batch.  Per pair the coordinates of the thread's <= 4 samples are staged once (cp.async into a
// thread-private shared-memory area, prefetched one pair ahead), turned into the pair geometry
// (unit vector, r^-3) held in registers, and ALL series of the pair are produced from it:
//   pass 0, per item (p, q):  Y21_p Y22_p F21_p F22_p | Y21_q Y22_q F21_q F22_q | Y20_p+iY20_q | F20_p+iF20_q
//                             (the real words of p wait in a shared-memory scratch for q) and the
//                             per-pair sums of r^-3 are accumulated;
//   pass 1, per item:         (r^-3_p - mean_p) + i (r^-3_q - mean_q)      (needs the sums of pass 0)
// Each series is an M1-point column FFT (half-zero first radix, radix-8 second step) written as
// T[item*11 + slot][k1][n2] for k_rows (slot table in kernels.h).
struct ColsDDParams {
    const double* X;
    const int* pi;
    const int* pj;
    i64 pair0;       // first pair of the batch
    i64 n_pairs;     // total pairs (pairs >= n_pairs do not exist)
    i64 items;       // pair-pairs in the batch
    i64 N;
    double a, b, c;
    cd* T;
    i64 M;
    double inv_M;
    double* rsum;    // [2*items] sums of r^-3 of the batch's pairs
    const cd* twp;
    int pass;        // 0: slots != 5, 1: slot 5
};

template <int M1>
__global__ void __launch_bounds__(256, 2) k_cols_dd(const ColsDDParams p) {
    constexpr int R1 = Plan8<M1>::R1;
    constexpr int R2 = 8;
    constexpr int NB = M1 / 8;
    constexpr int C = 256 / NB;
    constexpr int P1 = 8;
    constexpr int NFR = R2 / 2;
    constexpr int QH = (R1 + 1) / 2;
    extern __shared__ double2 smem[];
    __shared__ double sred[16];
    const int tid = threadIdx.x;
    const int c = tid % C;
    const int bb = tid / C;
    const int n2 = blockIdx.x * C + c;
    double* stage = reinterpret_cast<double*>(smem + pad16(M1) * C);   // [24][256] staged coordinates
    double* scratch = stage + 24 * 256;                                  // [3*NFR][256] real words of pair p
    cd tw[R2];
    {
        const cd base = w_exact(((i64)n2 * bb) % p.M, p.inv_M);
        const cd stp = w_exact(((i64)n2 * R1) % p.M, p.inv_M);
        tw[0] = base;
        tw[1] = stp;
#pragma unroll
        for (int j = 2; j < R2; ++j) tw[j] = (j & 1) ? cmul(tw[j - 1], stp) : csq(tw[j >> 1]);
#pragma unroll
        for (int j = 1; j < R2; ++j) tw[j] = cmul(tw[j], base);
    }
    auto frame_of = [&](int fs) -> i64 {
        const int it = fs / QH, q = fs % QH;
        return (i64)(q * P1 + bb + it * NB) * 4096 + n2;
    };
    // stage the coordinates of local pair index lp (= 2*item + h) of this CTA's sequence
    auto prefetch = [&](i64 lp) {
        const i64 pr = p.pair0 + lp;
        if (lp < 2 * p.items && pr < p.n_pairs) {
            const double* xi = p.X + (i64)p.pi[pr] * 3 * p.N;
            const double* xj = p.X + (i64)p.pj[pr] * 3 * p.N;
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) {
                const i64 frame = frame_of(fs);
                if (frame < p.N) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        cp_async8(stage + (fs * 6 + d) * 256 + tid, xi + d * p.N + frame);
                        cp_async8(stage + (fs * 6 + 3 + d) * 256 + tid, xj + d * p.N + frame);
                    }
                }
            }
        }
        cp_async_commit();
    };
    // one series: samples -> M1-point column FFT -> tile of T
    auto emit = [&](const cd (&smp)[NFR], i64 item, int slot) {
#pragma unroll
        for (int it = 0; it < R2 / R1; ++it) {
            const int m = bb + it * NB;
            cd u[R1];
#pragma unroll
            for (int q = 0; q < R1; ++q) u[q] = (q < QH) ? smp[it * QH + q] : make_double2(0.0, 0.0);
            if (R1 >= 2) fft_dif_halfzero<R1>(u);
            apply_powers<R1>(u, ldg_cd(p.twp + m * (4096 / M1)));
#pragma unroll
            for (int kq = 0; kq < R1; ++kq) smem[pad16(kq * P1 + m) * C + c] = u[brev(kq, R1)];
        }
        __syncthreads();
        cd u[R2];
#pragma unroll
        for (int q = 0; q < R2; ++q) u[q] = smem[pad16(bb * P1 + q) * C + c];
        fft_dif<R2>(u);
        cd* dst = p.T + (item * 11 + slot) * p.M + n2;
#pragma unroll
        for (int j = 0; j < R2; ++j) dst[(i64)(bb + R1 * j) * 4096] = cmul(u[brev(j, R2)], tw[j]);
        __syncthreads();
    };

    const double C20 = 0.31539156525252000603, C21 = -0.77254840404638362, C22 = 0.38627420202319181;
    const double KF = 1.5853309190424043;
    prefetch(2 * (i64)blockIdx.y);
    for (i64 item = blockIdx.y; item < p.items; item += gridDim.y) {
        double racc[2] = {0.0, 0.0};
        double rq[NFR];  // pass 1: r^-3 - mean of pair p, kept while pair q is processed
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const i64 pr = p.pair0 + 2 * item + h;
            const bool exists = pr < p.n_pairs;
            // geometry of this pair at the thread's samples: unit vector (x,y,z)/r and r^-3
            double gx[NFR], gy[NFR], gz[NFR], g3[NFR];
            cp_async_wait_all();
#pragma unroll
            for (int fs = 0; fs < NFR; ++fs) {
                gx[fs] = gy[fs] = gz[fs] = g3[fs] = 0.0;
                if (exists && frame_of(fs) < p.N) {
                    double v[6];
#pragma unroll
                    for (int d = 0; d < 6; ++d) v[d] = stage[(fs * 6 + d) * 256 + tid];
                    double dx, dy, dz, r2;
                    pair_geom(v, p.a, p.b, p.c, dx, dy, dz, r2);
                    double rinv = rsqrt(r2);
                    rinv = rinv * fma(-0.5 * r2, rinv * rinv, 1.5);
                    gx[fs] = dx * rinv; gy[fs] = dy * rinv; gz[fs] = dz * rinv;
                    g3[fs] = rinv * rinv * rinv;
                }
            }
            // next pair of this CTA's sequence: q of this item, or p of the next item
            prefetch(h == 0 ? 2 * item + 1 : 2 * (item + gridDim.y));
            cd smp[NFR];
            if (p.pass == 0) {
                const int so = h;  // slot offset of the complex series of this pair
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) smp[fs] = make_double2(C21 * gz[fs] * gx[fs], C21 * gz[fs] * gy[fs]);
                emit(smp, item, 1 + so);
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs)
                    smp[fs] = make_double2(C22 * (gx[fs] * gx[fs] - gy[fs] * gy[fs]), C22 * 2.0 * gx[fs] * gy[fs]);
                emit(smp, item, 3 + so);
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) {
                    const double k = KF * g3[fs] * C21 * gz[fs];
                    smp[fs] = make_double2(k * gx[fs], k * gy[fs]);
                }
                emit(smp, item, 7 + so);
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) {
                    const double k = KF * g3[fs] * C22;
                    smp[fs] = make_double2(k * (gx[fs] * gx[fs] - gy[fs] * gy[fs]), k * 2.0 * gx[fs] * gy[fs]);
                }
                emit(smp, item, 9 + so);
#pragma unroll
                for (int fs = 0; fs < NFR; ++fs) racc[h] += g3[fs];
                if (h == 0) {  // real words of p wait for q
#pragma unroll
                    for (int fs = 0; fs < NFR; ++fs) {
                        const double y20 = (exists && frame_of(fs) < p.N) ? C20 * (3.0 * gz[fs] * gz[fs] - 1.0) : 0.0;
                        scratch[(fs * 2 + 0) * 256 + tid] = y20;
                        scratch[(fs * 2 + 1) * 256 + tid] = KF * g3[fs] * y20;
                    }
                } else {
#pragma unroll
                    for (int fs = 0; fs < NFR; ++fs) {
                        const double y20 = (exists && frame_of(fs) < p.N) ? C20 * (3.0 * gz[fs] * gz[fs] - 1.0) : 0.0;
                        smp[fs] = make_double2(scratch[(fs * 2 + 0) * 256 + tid], y20);
                    }
                    emit(smp, item, 0);
#pragma unroll
                    for (int fs = 0; fs < NFR; ++fs) {
                        const double y20 = (exists && frame_of(fs) < p.N) ? C20 * (3.0 * gz[fs] * gz[fs] - 1.0) : 0.0;
                        smp[fs] = make_double2(scratch[(fs * 2 + 1) * 256 + tid], KF * g3[fs] * y20);
                    }
                    emit(smp, item, 6);
                }
            } else {
                const double mean = exists ? p.rsum[2 * item + h] / (double)p.N : 0.0;
                if (h == 0) {
#pragma unroll
                    for (int fs = 0; fs < NFR; ++fs) rq[fs] = (frame_of(fs) < p.N) ? g3[fs] - mean : 0.0;
                } else {
#pragma unroll
                    for (int fs = 0; fs < NFR; ++fs)
                        smp[fs] = make_double2(rq[fs], (exists && frame_of(fs) < p.N) ? g3[fs] - mean : 0.0);
                    emit(smp, item, 5);
                }
            }
        }
        if (p.pass == 0) {  // per-pair sums of r^-3 over this tile's samples
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                racc[0] += __shfl_down_sync(0xffffffffu, racc[0], o);
                racc[1] += __shfl_down_sync(0xffffffffu, racc[1], o);
            }
            if ((tid & 31) == 0) { sred[2 * (tid >> 5)] = racc[0]; sred[2 * (tid >> 5) + 1] = racc[1]; }
            __syncthreads();
            if (tid == 0) {
                double ta = 0.0, tb = 0.0;
                for (int w8 = 0; w8 < 8; ++w8) { ta += sred[2 * w8]; tb += sred[2 * w8 + 1]; }
                atomicAdd(p.rsum + 2 * item, ta);
                atomicAdd(p.rsum + 2 * item + 1, tb);
            }
        }
    }
}

}  // namespace dynpy
