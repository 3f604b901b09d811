// dynpy_b200 — second-pass (row) kernel of the four-step transform, accumulate-power flavour.
//
// One work unit = one row of 4096 complex points T[fft][k1][0..4096) (written by a column
// kernel, inter-pass twiddle already applied).  The unit list is ordered (slot, k1, item) and cut
// into equal contiguous ranges, one per CTA, so every CTA of the grid (2 per SM) carries the same
// load and stays on one accumulator row for long stretches: |X|^2 is summed in registers and
// flushed (coalesced fp64 REDs) only when (slot, k1) changes.
//
// Row transform 4096 = 16 x 16 x 16, decimation in frequency, 16 points per thread:
//   step 1  thread t: x[t + 256 q] -> fft16 -> times W_256^{(t/16) k}   (half-warp-uniform table look-ups;
//           the other factor W_4096^{(t%16) k} of the step twiddle commutes with step 2 and is merged into
//           its twiddle) -> shared memory
//   step 2  thread (h = t/16, m = t%16): fft16 over q' -> times W_4096^{m (h + 16 k2)} (one chain of
//           complex products from two table entries) -> shared memory (same half warp reads it back)
//   step 3  fft16 -> |X|^2 += into 16 registers; X index = t/16 + 16 (t%16) + 256 j.
// The next row is pulled into L2 by one bulk prefetch while the current row is transformed.
// fp64 instructions per point: 3*148/16 (codelets) + 4*15/16 + 4 + 4*15/16 + 2 = 41.3.
//
// The same row transform also serves the write epilogues of the generic engine (template parameter MODE): per-series
// power spectra (staged through a second shared-memory region, stored as 256-byte runs) and the real outputs of the
// inverse transforms (k = k1 + M1 k2, rows dealt round-robin over the CTAs).
#pragma once
#include "codelets.cuh"
#include "fft_engine.cuh"

namespace dynpy {

// Layout of the intermediate T: row k1 of transform t starts at (t * M1 + k1) * DYNPY_T_RS (cd units).  The
// row stride is 4096 + a small pad so that the hundreds of rows streamed concurrently (one per CTA, all
// a multiple of 64 KB apart otherwise) do not collide on the same HBM channels / banks.
#ifndef DYNPY_T_RS
#define DYNPY_T_RS (4096 + 16)
#endif
// 1: the next row is loaded into 64 registers under step 3 of the current one; 0 (A/B): every row is read at the
// top of its own iteration (out of L2, where the bulk prefetch put it), which frees those registers
#ifndef DYNPY_ROWS_NX
#define DYNPY_ROWS_NX 1
#endif

// timing experiment only (results are garbage): the T addresses of all items folded onto DYNPY_T_ALIAS items, so that
// T stays resident in L2 — measures what the two transform kernels would gain if T never went to DRAM
#ifdef DYNPY_T_ALIAS
#define T_ITEM(it) ((it) % (i64)DYNPY_T_ALIAS)
#else
#define T_ITEM(it) (it)
#endif

struct Rows2Params {
    const cd* T;            // [items][nslots][M1][rs]
    double* S;              // [nacc][M] accumulated |X|^2, internal order k1*4096 + k2
    i64 M;
    int M1;
    int rs;                 // row stride of T in cd units (4096, or DYNPY_T_RS for the padded layout)
    int nslots;
    i64 items;
    unsigned long long acc_pack;  // 4-bit accumulator id per slot
    int nsel;                     // slots handled by this launch ...
    unsigned long long sel_pack;  // ... and their ids, 4 bits each
    const cd* twp;          // W_4096^j
    // optional mean removal of one slot:  x <- x - (mu_re + i mu_im) * R,  mu = rsum[2 item (+1)] * inv_n
    const cd* R;            // [M1][rs] column transform of the all-ones series (nullptr: off)
    const double* rsum;     // [2 * items]
    int fix_slot;
    double inv_n;
    int pf_dist;            // rows of look-ahead of the L2 bulk prefetch (0: off)
    // write modes (MODE != 0): T holds transforms f_base .. f_base + items - 1 (nslots = 1), units ordered (f, k1)
    double* out;            // WRITE_POWER: P[f][M] internal order; WRITE_REAL(2): out[f (2f, 2f+1)][n_out], k = k1 + M1 k2
    i64 f_base;
    i64 n_out;
    i64 n_series;           // WRITE_REAL2: second output row exists while 2 f + 1 < n_series
    double scale;
};
// epilogues of the row kernel (the values of fft_engine.cuh's EpiMode)
enum Rows2Mode { ROWS2_ACC_POWER = 0, ROWS2_WRITE_REAL = 1, ROWS2_WRITE_POWER = 2, ROWS2_WRITE_REAL2 = 3,
                 ROWS2_WRITE_INPLACE = 4 };   // 4: the transformed row goes back into its own T row (complex, k2 order)

__device__ __forceinline__ cd ldcg_cd2(const cd* p) {
    double2 r;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

template <bool FIX, int MODE = ROWS2_ACC_POWER>
__global__ void __launch_bounds__(256, 2) k_rows2(const Rows2Params p) {
    extern __shared__ double2 smem[];
    const int tid = threadIdx.x;
    const i64 units = (MODE == ROWS2_ACC_POWER ? (i64)p.nsel : (i64)1) * p.M1 * p.items;
    // accumulate / power modes: equal contiguous ranges; real-output modes: rows dealt round-robin (ustep = grid size), so
    // that at any moment the CTAs of the grid hold CONSECUTIVE rows k1 of the same few transforms — their 8-byte
    // outputs k = k1 + M1 k2 share 32-byte sectors, which are then completed in L2 within one row time (with
    // contiguous ranges a sector's four writers were microseconds apart: 3.5 x the DRAM write traffic, read-fills)
    constexpr bool RR = MODE == ROWS2_WRITE_REAL || MODE == ROWS2_WRITE_REAL2;
    const i64 ustep = RR ? (i64)gridDim.x : 1;
    const i64 u0 = RR ? (i64)blockIdx.x : units * blockIdx.x / gridDim.x;
    const i64 u1 = RR ? units : units * (blockIdx.x + 1) / gridDim.x;
    if (u0 >= u1) return;

    const int h = tid >> 4, m = tid & 15;
    // padded shared-memory indices (pad16(i) = i + i / 16) written out as base + constant: pad16(tid + 256 k) =
    // pad16(tid) + 272 k, pad16(256 h + 16 q + m) = 272 h + m + 17 q, pad16(16 tid + q) = 17 tid + q
    const int b1 = tid + h, b2 = 272 * h + m, b3 = 17 * tid;
    // step-2 merged twiddle chain: c0 = W_4096^{m h}, step = W_256^m
    const cd c0 = ldg_cd(p.twp + m * h);
    const cd cs = ldg_cd(p.twp + 16 * m);
    // step-1 twiddles W_256^{h k} = twp[16 h k] in shared memory, [k][h]: the two h of a warp are adjacent, so a
    // look-up costs one wavefront (from global they miss the small L1 and stall on L2 latency)
    cd* stw1 = smem + pad16(4096);
    stw1[tid] = ldg_cd(p.twp + 16 * (tid & 15) * (tid >> 4));
    __syncthreads();

    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0;

    // position in the unit list (slot index, k1, item; item fastest): decoded once by division, then advanced by
    // increments (the two 64-bit divisions per row cost ~240 instructions between steps 1 and 2)
    struct Cursor { int si, k1; i64 item; };
    auto decode = [&](i64 u) -> Cursor {
        Cursor c;
        if constexpr (MODE != ROWS2_ACC_POWER) {   // (f, k1), k1 fastest: the rows of one transform stay together
            c.si = 0;
            c.item = u / p.M1;
            c.k1 = (int)(u - c.item * p.M1);
            return c;
        }
        const i64 sk = u / p.items;
        c.item = u - sk * p.items;
        c.si = (int)(sk / p.M1);
        c.k1 = (int)(sk - (i64)c.si * p.M1);
        return c;
    };
    const int rr_dk = (int)(ustep % p.M1);
    const i64 rr_di = ustep / p.M1;
    auto advance = [&](Cursor& c) {
        if constexpr (RR) {
            c.k1 += rr_dk; c.item += rr_di;
            if (c.k1 >= p.M1) { c.k1 -= p.M1; ++c.item; }
            return;
        }
        if constexpr (MODE != ROWS2_ACC_POWER) {
            if (++c.k1 == p.M1) { c.k1 = 0; ++c.item; }
            return;
        }
        if (++c.item == p.items) {
            c.item = 0;
            if (++c.k1 == p.M1) { c.k1 = 0; ++c.si; }
        }
    };
    auto slot_of = [&](const Cursor& c) -> int {
        if constexpr (MODE != ROWS2_ACC_POWER) return 0;
        return (int)((p.sel_pack >> (4 * c.si)) & 15ull);
    };
    auto ptr_of = [&](const Cursor& c) -> const cd* {
        return p.T + ((T_ITEM(c.item) * p.nslots + slot_of(c)) * p.M1 + c.k1) * (i64)p.rs;
    };
    auto flush = [&](int slot, int k1) {
        double* sr = reinterpret_cast<double*>(smem);
        const int kk0 = h + 16 * m;
#pragma unroll
        for (int j = 0; j < 16; ++j) sr[kk0 + 256 * j] = acc[j];
        __syncthreads();
        double* dst = p.S + (i64)((p.acc_pack >> (4 * slot)) & 15ull) * p.M + ((i64)k1 << 12);
        for (int k = tid; k < 4096; k += 256) atomicAdd(dst + k, sr[k]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0;
    };

    // element n2 = tid + 256 q of a row: tile (tid / TILE) + (256 / TILE) q, offset tid % TILE
    auto t_off = [&](int q) -> i64 { return tid + 256 * q; };
    Cursor cur = decode(u0);
    const i64 pf_ahead = (p.pf_dist > 0 ? p.pf_dist : 0) * ustep;
    Cursor pfc = decode(u0 + pf_ahead < u1 ? u0 + pf_ahead : u0);   // prefetch cursor
    int slot = slot_of(cur), k1 = cur.k1;
    i64 item = cur.item;
    const cd* src = ptr_of(cur);

#if DYNPY_ROWS_NX
    cd nx[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) nx[q] = ldcg_cd2(src + t_off(q));
#endif

    for (i64 u = u0; u < u1; u += ustep) {
        cd v[16];
#if DYNPY_ROWS_NX
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = nx[q];
#else
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = ldcg_cd2(src + t_off(q));
#endif
        if constexpr (FIX) {
            const double mr = p.rsum[2 * item] * p.inv_n, mi = p.rsum[2 * item + 1] * p.inv_n;
            const cd* rr = p.R + (i64)k1 * p.rs;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cd r = ldg_cd(rr + t_off(q));
                v[q].x -= fma(mr, r.x, -(mi * r.y));
                v[q].y -= fma(mr, r.y, mi * r.x);
            }
        }
        // ---- step 1
        fft16(v);
        smem[b1] = v[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) smem[b1 + 272 * k] = cmul(v[k], stw1[k * 16 + h]);
        __syncthreads();
        // next row: decode now (its registers are loaded under step 3); the row after it is pulled into
        // L2 by one bulk prefetch per chunk (DRAM latency under this load is several microseconds)
        int nslot = slot, nk1 = k1;
        i64 nitem = item;
        if (u + ustep < u1) {
            advance(cur);
            src = ptr_of(cur);
            nslot = slot_of(cur); nk1 = cur.k1; nitem = cur.item;
        }
        if (p.pf_dist > 0 && u + pf_ahead < u1) {
            if (tid == 0) {
                const cd* pf = ptr_of(pfc);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf), "r"(4096 * 16) : "memory");
            }
            advance(pfc);
        }
        // ---- step 2
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = smem[b2 + 17 * q];
        fft16(v);
        {
            cd w = c0;
            asm volatile("" : "+d"(w.x), "+d"(w.y));  // keep the chain inside the loop (15 hoisted values would spill)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                smem[b2 + 17 * k] = cmul(v[k], w);
                if (k < 15) w = cmul(w, cs);
            }
        }
        // next row (already on its way to L2) -> registers: issued as soon as the step-2 values have left the
        // registers, so that the loads fly under the step-3 reads, the barrier and the last codelet;
        // unconditional so that nx is dead during steps 1 and 2 (the last iteration re-reads its own row)
        __syncwarp();
        // ---- step 3
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = smem[b3 + q];
#if DYNPY_ROWS_NX
#pragma unroll
        for (int q = 0; q < 16; ++q) nx[q] = ldcg_cd2(src + t_off(q));
#endif
        __syncthreads();  // every thread has its step-3 inputs: the buffer may be rewritten
        fft16(v);
        if constexpr (MODE == ROWS2_ACC_POWER) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fma(v[j].x, v[j].x, fma(v[j].y, v[j].y, acc[j]));
            if (u + ustep >= u1 || nslot != slot || nk1 != k1) flush(slot, k1);
        } else if constexpr (MODE == ROWS2_WRITE_POWER) {
            // |X|^2 of the row, k2 = h + 16 m + 256 j, staged through a second shared-memory region (padded: k + k / 16)
            // so that the row leaves as 256-byte runs; the region is rewritten only after the next row's two barriers
            double* sr = reinterpret_cast<double*>(smem + pad16(4096) + 256);
            const int kk0 = h + 17 * m;
#pragma unroll
            for (int j = 0; j < 16; ++j) sr[kk0 + 272 * j] = fma(v[j].x, v[j].x, v[j].y * v[j].y);
            __syncthreads();
            double* dst = p.out + (p.f_base + item) * p.M + ((i64)k1 << 12);
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[tid + 256 * i] = sr[tid + (tid >> 4) + 272 * i];
        } else if constexpr (MODE == ROWS2_WRITE_INPLACE) {
            // X[k2], k2 = h + 16 m + 256 j, back into the row it came from (every row belongs to one CTA, and it was
            // read completely before step 1), staged through the exchange buffer — free since the barrier above — so
            // that the row leaves as 512-byte runs; k_rows_to_natural then transposes (k1, k2) -> k = k1 + M1 k2 with
            // both sides coalesced.  The scattered form (MODE 1 / 3) costs 16 - 32 stores of 8 bytes per thread into
            // separate sectors: 0.56 ms against 0.25 ms for the power epilogue on the same launch.
            const int kk0 = h + 17 * m;
#pragma unroll
            for (int j = 0; j < 16; ++j) smem[kk0 + 272 * j] = v[j];
            __syncthreads();
            cd* dst = const_cast<cd*>(p.T) + (item * p.M1 + k1) * (i64)p.rs;
            // only k = k1 + M1 k2 < n_out is ever read back: k2 < ceil((n_out - k1) / M1)
            const i64 k2_end = p.n_out > k1 ? (p.n_out - k1 + p.M1 - 1) / p.M1 : 0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (256 * i < k2_end) dst[tid + 256 * i] = smem[tid + (tid >> 4) + 272 * i];
            __syncthreads();   // the buffer is rewritten by step 1 of the next row
        } else {
            // natural index k = k1 + M1 k2: consecutive k1 (consecutive CTAs of the grid, see RR above) fill the same
            // 32-byte sectors at about the same time
            const i64 f = p.f_base + item;
            const bool two = MODE == ROWS2_WRITE_REAL2 && 2 * f + 1 < p.n_series;
            double* o0 = p.out + (MODE == ROWS2_WRITE_REAL2 ? 2 * f : f) * p.n_out;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const i64 k = (i64)k1 + (i64)p.M1 * (h + 16 * m + 256 * j);
                if (k < p.n_out) {
                    o0[k] = v[j].x * p.scale;
                    if (two) o0[p.n_out + k] = v[j].y * p.scale;
                }
            }
        }
        slot = nslot; k1 = nk1; item = nitem;
    }
}

}  // namespace dynpy
