// Host launcher of the persistent fused dipolar kernel (dd_stream.cuh).
#include <stdlib.h>

#include "dd_stream.cuh"
#include "kernels.h"

namespace dynpy {

bool dd_stream_supported(i64 M, i64 n_pairs) {
    const char* e = getenv("DYNPY_DD_STREAM");  // experimental persistent kernel: opt-in
    if (!e || e[0] != '1') return false;
    if (M != 4096 * 16 && M != 4096 * 32 && M != 4096 * 64) return false;
    return n_pairs >= 64;  // the stream needs far more transforms than ring slots
}

i64 dd_stream_workspace_bytes(i64 M, i64 n_pairs) {
    i64 W = ((i64)40 << 20) / (M * (i64)sizeof(cd));
    if (W < 4) W = 4;
    return W * M * (i64)sizeof(cd) + n_pairs * 8 + 4096 + 8 * W * 2;
}

template <int M1>
static cudaError_t launch_stream_M1(StreamParams& p, cudaStream_t st) {
    auto kern = k_dd_stream<M1>;
    constexpr int NBc = M1 / 8, Cc = 256 / NBc;
    const size_t prod = (size_t)pad16(M1) * Cc * sizeof(cd) + (size_t)24 * 256 * sizeof(double);
    const size_t cons = (size_t)pad16(4096) * sizeof(cd);
    const size_t smem = prod > cons ? prod : cons;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
    if (e != cudaSuccess) return e;
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    constexpr int NB = M1 / 8, C = 256 / NB, NPOS = 4096 / C;
    const int want = 256;  // 128 producers + 128 consumers
    if (per_sm * n_sm < want) return cudaErrorCooperativeLaunchTooLarge;
    p.n_prod = 128;
    p.kp = 128 / NPOS;
    p.kc = 128 / M1;
    void* args[] = {(void*)&p};
    ProfScope prof(PROF_ROWS, st);
    return cudaLaunchCooperativeKernel((void*)kern, dim3(want), dim3(256), args, smem, st);
}

// spectrum[7][M] += ...; ws carved as: ring | counters | rsum | err.  Synchronises the stream to
// read the error flag (the call runs for seconds; one sync is irrelevant).
int dd_stream_run(const double* X, i64 N, const int* pi, const int* pj, i64 P, double a, double b, double c,
                  double* S, i64 M, void* ws, i64 ws_bytes, const cd* tw, cudaStream_t st) {
    if (ws_bytes < dd_stream_workspace_bytes(M, P)) {
        set_error("dd stream: workspace too small (need %lld bytes)", (long long)dd_stream_workspace_bytes(M, P));
        return DYNPY_ERR_WORKSPACE;
    }
    i64 W = ((i64)40 << 20) / (M * (i64)sizeof(cd));
    if (W < 4) W = 4;
    char* base = (char*)ws;
    StreamParams p;
    memset(&p, 0, sizeof(p));
    p.T = (cd*)base;
    i64 off = W * M * (i64)sizeof(cd);
    p.ready = (unsigned long long*)(base + off);
    p.done = p.ready + W;
    p.total = p.done + W;
    p.err = (int*)(p.total + 1);
    i64 small = (2 * W + 2) * 8;
    small = (small + 255) / 256 * 256;
    p.rsum = (double*)(base + off + small);
    DYNPY_CHECK(cudaMemsetAsync(base + off, 0, small + P * 8, st), "dd stream memset");
    p.X = X; p.pi = pi; p.pj = pj; p.n_pairs = P; p.N = N; p.a = a; p.b = b; p.c = c;
    p.S = S; p.M = M; p.inv_M = 1.0 / (double)M; p.W = (int)W; p.twp = tw;
    const i64 I = (P + 1) / 2;
    const i64 cnt[7] = {I, P, P, I, I, P, P};
    p.bnd[0] = 0;
    for (int t = 0; t < 7; ++t) p.bnd[t + 1] = p.bnd[t] + cnt[t];
    p.spin_limit = 20000000;
    long long* dbg = nullptr;
    if (getenv("DYNPY_DD_STREAM_DEBUG")) {
        cudaMalloc(&dbg, 4 * 512 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 4 * 512 * sizeof(long long), st);
        p.dbg = dbg;
    }
    const int M1 = (int)(M / 4096);
    cudaError_t e = M1 == 16 ? launch_stream_M1<16>(p, st) : M1 == 32 ? launch_stream_M1<32>(p, st)
                                                                       : launch_stream_M1<64>(p, st);
    if (e != cudaSuccess) return fail_cuda(e, "dd stream launch");
    int herr = 0;
    DYNPY_CHECK(cudaMemcpyAsync(&herr, p.err, sizeof(int), cudaMemcpyDeviceToHost, st), "dd stream flag");
    DYNPY_CHECK(cudaStreamSynchronize(st), "dd stream sync");
    if (dbg) {
        static long long h[4 * 512];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        for (int role = 0; role < 2; ++role) {
            double w = 0, t = 0, n = 0; int cnt = 0;
            for (int b = role * 128; b < role * 128 + 128; ++b) { w += h[4*b]; t += h[4*b+2]; n += h[4*b+3]; ++cnt; }
            fprintf(stderr, "[dd_stream] %s: iterations/CTA %.0f  cycles/iter %.0f  of which waiting %.0f\n",
                    role ? "consumers" : "producers", n / cnt, t / n, w / n);
        }
    }
    if (herr) {
        set_error("dd stream: a producer/consumer hand-over timed out (kernel aborted, results invalid)");
        return DYNPY_ERR_LAUNCH;
    }
    return DYNPY_OK;
}

// ---- batched fused column pass (k_cols_dd) -------------------------------------------------------
bool dd_cols_supported(i64 M) {
    const char* e = getenv("DYNPY_DD_FUSED_COLS");
    if (e && e[0] == '0') return false;
    return M == 4096 * 16 || M == 4096 * 32 || M == 4096 * 64;
}

template <int M1>
static cudaError_t launch_cols_dd_M1(ColsDDParams& p, int n_sm, cudaStream_t st) {
    auto kern = k_cols_dd<M1>;
    constexpr int NB = M1 / 8, C = 256 / NB, NPOS = 4096 / C;
    const size_t smem = (size_t)pad16(M1) * C * sizeof(cd) + (size_t)(24 + 8) * 256 * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (int pass = 0; pass < 2; ++pass) {
        p.pass = pass;
        // one resident wave: every CTA (2 per SM) loops over its share of the items
        int gy = (2 * n_sm) / NPOS;
        if (gy > p.items) gy = (int)p.items;
        if (gy < 1) gy = 1;
        ProfScope prof(PROF_COLS, st);
        kern<<<dim3(NPOS, gy), 256, smem, st>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_cols_dd(const double* X, i64 N, const int* pi, const int* pj, i64 pair0, i64 n_pairs, i64 items,
                           double a, double b, double c, cd* T, i64 M, double* rsum, const cd* tw, int n_sm,
                           cudaStream_t st) {
    ColsDDParams p;
    memset(&p, 0, sizeof(p));
    p.X = X; p.pi = pi; p.pj = pj; p.pair0 = pair0; p.n_pairs = n_pairs; p.items = items; p.N = N;
    p.a = a; p.b = b; p.c = c; p.T = T; p.M = M; p.inv_M = 1.0 / (double)M; p.rsum = rsum; p.twp = tw;
    cudaError_t e = cudaMemsetAsync(rsum, 0, sizeof(double) * 2 * items, st);
    if (e != cudaSuccess) return e;
    const int M1 = (int)(M / 4096);
    return M1 == 16 ? launch_cols_dd_M1<16>(p, n_sm, st) : M1 == 32 ? launch_cols_dd_M1<32>(p, n_sm, st)
                                                                    : launch_cols_dd_M1<64>(p, n_sm, st);
}

}  // namespace dynpy
