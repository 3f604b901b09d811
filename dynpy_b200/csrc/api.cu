// dynpy_b200 — the C-ABI (include/dynpy_b200.h): argument checks, workspace carving,
// batch loops.  All compute is in the kernels of fft_engine.cuh / kernels_misc.cu.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "../../include/dynpy_b200.h"
#include "kernels.h"

namespace dynpy {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail_cuda(cudaError_t e, const char* where) {
    set_error("%s: %s", where, cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky-less error
    return DYNPY_ERR_LAUNCH;
}

static DeviceState g_dev[64];
static std::mutex g_dev_mutex;   // first use of a device (twiddle table) and the profiling lists

// ---- profiling hooks -------------------------------------------------------------------
struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> ev[PROF_NCLS];  // begin/end pairs
    std::vector<cudaEvent_t> pool;
};
static ProfState g_prof;
static cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_prof.pool.empty()) { e = g_prof.pool.back(); g_prof.pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int cls, cudaStream_t st) {
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    g_prof.ev[cls].push_back(e);
}
void prof_end(int cls, cudaStream_t st) {
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    g_prof.ev[cls].push_back(e);
}

int get_device_state(DeviceState** out, cudaStream_t st) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64) {
        set_error("no CUDA device (%s)", e == cudaSuccess ? "bad ordinal" : cudaGetErrorString(e));
        return DYNPY_ERR_NO_DEVICE;
    }
    DeviceState& d = g_dev[dev];
    std::lock_guard<std::mutex> lock(g_dev_mutex);  // calls from several host threads may race on the first use
    if (!d.ready) {
        cudaDeviceProp prop;
        DYNPY_CHECK(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
        if (prop.major < 10) {
            set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
            return DYNPY_ERR_NO_DEVICE;
        }
        d.n_sm = prop.multiProcessorCount;
        DYNPY_CHECK(cudaMalloc(&d.tw4096, 4096 * sizeof(double2)), "cudaMalloc(twiddles)");
        DYNPY_CHECK(launch_init_twiddles(d.tw4096, st), "init twiddles");
        DYNPY_CHECK(cudaStreamSynchronize(st), "init twiddles sync");
        d.ready = true;
    }
    *out = &d;
    return DYNPY_OK;
}

static inline i64 align_up(i64 x, i64 a) { return (x + a - 1) / a * a; }
static inline bool pow2(i64 m) { return m > 0 && (m & (m - 1)) == 0; }

static int check_M(i64 M, i64 n) {
    if (!fft_length_ok(M)) {
        set_error("transform length M=%lld must be a power of two in [32, 2^24] or 4096 x {20, 24, 40, 50}", (long long)M);
        return DYNPY_ERR_INVALID;
    }
    if (n < 1 || 2 * n - 1 > M) {
        set_error("series length n=%lld needs M >= 2n-1 (M=%lld)", (long long)n, (long long)M);
        return DYNPY_ERR_INVALID;
    }
    return DYNPY_OK;
}

}  // namespace dynpy

using namespace dynpy;

extern "C" {
#pragma GCC visibility push(default)

int dynpy_version(void) { return 100; }
const char* dynpy_last_error(void) { return g_err; }

int dynpy_device_sm_count(void) {
    DeviceState* d;
    int rc = get_device_state(&d, 0);
    return rc ? rc : d->n_sm;
}

int dynpy_profile_enable(int on) {
    g_prof.on = on != 0;
    return DYNPY_OK;
}

// Sum of device time (ms) and number of launches per kernel class since the last read;
// synchronises the device.  ms[4], launches[4]: words, cols, rows, other.
int dynpy_profile_read(double* ms_host, int64_t* launches_host) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail_cuda(e, "profile sync");
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    for (int c = 0; c < PROF_NCLS; ++c) {
        double tot = 0.0;
        auto& v = g_prof.ev[c];
        for (size_t k = 0; k + 1 < v.size(); k += 2) {
            float t = 0.f;
            cudaEventElapsedTime(&t, v[k], v[k + 1]);
            tot += t;
        }
        if (ms_host) ms_host[c] = tot;
        if (launches_host) launches_host[c] = (int64_t)(v.size() / 2);
        for (auto ev : v) g_prof.pool.push_back(ev);
        v.clear();
    }
    return DYNPY_OK;
}

int64_t dynpy_fft_length(int64_t n) { return fft_length_for(n); }

int64_t dynpy_fft_workspace_bytes(int64_t M, int64_t n_fft) {
    if (M <= DYNPY_ROW_LEN) return 0;
    if (n_fft < 1) n_fft = 1;
    return n_fft * M * (int64_t)sizeof(cd);
}

// ------------------------------------------------------------------------------------------ WK
int dynpy_wk_acf_sum_f64(const double* re, const double* im, int64_t n_series, int64_t n, int64_t series_stride,
                         int pack_real_pairs, const double* scale, double* spectrum, int64_t M, void* ws,
                         int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_M(M, n);
    if (rc) return rc;
    if (n_series < 0 || !re || !spectrum || series_stride < n || (pack_real_pairs && im)) {
        set_error("dynpy_wk_acf_sum_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    if (pack_real_pairs && scale) {
        set_error("dynpy_wk_acf_sum_f64: per-series scale is not supported together with pack_real_pairs");
        return DYNPY_ERR_INVALID;
    }
    if (n_series == 0) return DYNPY_OK;
    DeviceState* d;
    rc = get_device_state(&d, st);
    if (rc) return rc;
    PlanarLoader ld;
    memset(&ld, 0, sizeof(ld));
    ld.re = re;
    ld.nslots = 1;
    ld.N = n;
    ld.scale = scale;
    ld.bias_slot = -1;
    i64 n_fft;
    if (pack_real_pairs) {
        ld.im = re + series_stride;
        ld.hi = 2 * series_stride;
        n_fft = (n_series + 1) / 2;
        ld.n_im_valid = n_series / 2;
    } else {
        ld.im = im;
        ld.hi = series_stride;
        n_fft = n_series;
        ld.n_im_valid = n_series;
    }
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = spectrum;
    ep.M = M;
    ep.nslots = 1;
    ep.acc_pack = 0;
    i64 T_cap = M > DYNPY_ROW_LEN ? ws_bytes / (M * (i64)sizeof(cd)) : 0;
    if (M > DYNPY_ROW_LEN && (T_cap < 1 || !ws)) {
        set_error("workspace too small: need >= %lld bytes", (long long)(M * (i64)sizeof(cd)));
        return DYNPY_ERR_WORKSPACE;
    }
    DYNPY_CHECK(forward_planar_acc(ld, ep, n_fft, M, 1, (cd*)ws, T_cap, d->tw4096, d->n_sm, st), "wk forward");
    return DYNPY_OK;
}

int dynpy_ifft_acf_f64(const double* spectrum, int n_acc, int64_t M, double* acf, int64_t n_out, double scale,
                       void* ws, int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!fft_length_ok(M) || n_acc < 1 || n_out < 1 || n_out > M || !spectrum || !acf) {
        set_error("dynpy_ifft_acf_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    SpectrumLoader ld;
    ld.S = spectrum;
    ld.M = M;
    ld.M1 = M > DYNPY_ROW_LEN ? (int)(M / DYNPY_ROW_LEN) : 1;
    ld.natural = 0;
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = acf;
    ep.M = M;
    ep.nslots = 1;
    ep.n_out = n_out;
    ep.scale = scale;
    i64 T_cap = M > DYNPY_ROW_LEN ? ws_bytes / (M * (i64)sizeof(cd)) : 0;
    if (M > DYNPY_ROW_LEN && (T_cap < 1 || !ws)) {
        set_error("workspace too small: need >= %lld bytes", (long long)(M * (i64)sizeof(cd)));
        return DYNPY_ERR_WORKSPACE;
    }
    DYNPY_CHECK(forward_spectrum_real(ld, ep, n_acc, M, 1, (cd*)ws, T_cap, d->tw4096, d->n_sm, st), "final transform");
    return DYNPY_OK;
}

int64_t dynpy_wk_each_workspace_bytes(int64_t M, int64_t n_series) {
    if (n_series < 1) n_series = 1;
    i64 t = M > DYNPY_ROW_LEN ? (n_series < 16 ? n_series : 16) * M * (i64)sizeof(cd) : 0;
    const i64 copies = M > DYNPY_ROW_LEN ? 2 : 1;   // two-pass lengths: the spectra also in natural order
    return copies * n_series * align_up(M * 8, 256) + t + 512;
}

int dynpy_wk_acf_each_f64(const double* re, const double* im, int64_t n_series, int64_t n, int64_t series_stride,
                          double* acf, int64_t M, void* ws, int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_M(M, n);
    if (rc) return rc;
    if (n_series < 0 || !re || !acf || series_stride < n || !ws) {
        set_error("dynpy_wk_acf_each_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    if (n_series == 0) return DYNPY_OK;
    DeviceState* d;
    rc = get_device_state(&d, st);
    if (rc) return rc;
    // carve: T (up to 16 transforms) | P[cap][M]
    const bool two = M > DYNPY_ROW_LEN;
    i64 cap = n_series;
    while (cap > 1 && dynpy_wk_each_workspace_bytes(M, cap) > ws_bytes) cap = (cap + 1) / 2;
    if (dynpy_wk_each_workspace_bytes(M, cap) > ws_bytes) {
        set_error("dynpy_wk_acf_each_f64: workspace too small (need >= %lld bytes)",
                  (long long)dynpy_wk_each_workspace_bytes(M, 1));
        return DYNPY_ERR_WORKSPACE;
    }
    const i64 T_cap = two ? (cap < 16 ? cap : 16) : 0;
    cd* T = two ? (cd*)ws : nullptr;
    double* P = (double*)((char*)ws + align_up(T_cap * M * (i64)sizeof(cd), 256));
    double* Pn = two ? P + cap * M : nullptr;
    const i64 pstride = align_up(M * 8, 256) / 8;
    if (pstride != M) { set_error("internal: unaligned spectrum stride"); return DYNPY_ERR_INVALID; }
    for (i64 s0 = 0; s0 < n_series; s0 += cap) {
        const i64 cnt = n_series - s0 < cap ? n_series - s0 : cap;
        PlanarLoader ld;
        memset(&ld, 0, sizeof(ld));
        ld.re = re + s0 * series_stride;
        ld.im = im ? im + s0 * series_stride : nullptr;
        ld.hi = series_stride;
        ld.nslots = 1;
        ld.N = n;
        ld.n_im_valid = cnt;
        ld.bias_slot = -1;
        EpiParams ep;
        memset(&ep, 0, sizeof(ep));
        ep.out = P;
        ep.M = M;
        ep.nslots = 1;
        DYNPY_CHECK(forward_planar_power(ld, ep, cnt, M, 1, T, T_cap, d->tw4096, d->n_sm, st), "wk each forward");
        SpectrumLoader sl;
        sl.S = P;
        sl.M = M;
        sl.M1 = two ? (int)(M / DYNPY_ROW_LEN) : 1;
        sl.natural = 0;
        if (two) {
            DYNPY_CHECK(launch_spectrum_to_natural(P, Pn, cnt, sl.M1, st), "wk each transpose");
            sl.S = Pn;
            sl.natural = 1;
        }
        EpiParams e2;
        memset(&e2, 0, sizeof(e2));
        e2.out = acf + s0 * n;
        e2.M = M;
        e2.nslots = 1;
        e2.n_out = n;
        e2.scale = 1.0 / ((double)M * (double)n);
        DYNPY_CHECK(forward_spectrum_real(sl, e2, cnt, M, 1, T, T_cap, d->tw4096, d->n_sm, st), "wk each inverse");
    }
    return DYNPY_OK;
}

// ------------------------------------------------------------------------------------------ QR
int dynpy_qr_efg_to_spectrum_f64(const double* efg, int64_t n_nuclei, int64_t n_frames, double* spectrum, int64_t M,
                                 void* ws, int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_M(M, n_frames);
    if (rc) return rc;
    if (!efg || !spectrum || n_nuclei < 0) {
        set_error("dynpy_qr_efg_to_spectrum_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    if (n_nuclei == 0) return DYNPY_OK;
    DeviceState* d;
    rc = get_device_state(&d, st);
    if (rc) return rc;
    QrLoader ld;
    ld.V = efg;
    ld.N = n_frames;
    ld.nnuc = n_nuclei;
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = spectrum;
    ep.M = M;
    ep.nslots = 5;
    ep.acc_pack = 0x21210ull;  // slots 0..4 -> acc 0,1,2,1,2
    const i64 n_items = (n_nuclei + 1) / 2;
    i64 T_cap = M > DYNPY_ROW_LEN ? ws_bytes / (M * (i64)sizeof(cd)) : 0;
    if (M > DYNPY_ROW_LEN && (T_cap < 5 || !ws)) {
        set_error("workspace too small: need >= %lld bytes", (long long)(5 * M * (i64)sizeof(cd)));
        return DYNPY_ERR_WORKSPACE;
    }
    DYNPY_CHECK(forward_qr_acc(ld, ep, n_items * 5, M, 5, (cd*)ws, T_cap, d->tw4096, d->n_sm, st), "qr forward");
    return DYNPY_OK;
}

// ------------------------------------------------------------------------------------------ DD
int dynpy_wrap_f64(const double* x, double* out, int64_t n, double L, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !out || n < 0 || !(L > 0.0)) {
        set_error("dynpy_wrap_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_wrap(x, out, n, L, d->n_sm, st), "wrap");
    return DYNPY_OK;
}

int64_t dynpy_pdist_workspace_bytes(int64_t n) {
    int64_t npairs = n * (n - 1) / 2;
    return ((npairs + 255) / 256 + 1) * (int64_t)sizeof(int64_t);
}

int dynpy_pdist_ortho_f64(const double* ux, const double* uy, const double* uz, int64_t stride, int64_t n, double a,
                          double b, double c, const int64_t* index, double dmax, double* dx, double* dy, double* dz,
                          double* dr, int64_t* atom0, int64_t* atom1, int64_t* projection, int64_t* count, void* ws,
                          int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!count || n < 0 || stride < 1 || (n >= 2 && (!ux || !uy || !uz || !index))) {
        set_error("dynpy_pdist_ortho_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    const bool count_only = !dx && !dy && !dz && !dr && !atom0 && !atom1 && !projection;
    if (n >= 2 && !count_only && (!dx || !dy || !dz || !dr || !atom0 || !atom1 || !projection)) {
        set_error("dynpy_pdist_ortho_f64: output arrays missing (pass all seven, or none for a count-only call)");
        return DYNPY_ERR_INVALID;
    }
    if (n >= 2 && (ws_bytes < dynpy_pdist_workspace_bytes(n) || !ws)) {
        set_error("dynpy_pdist_ortho_f64: workspace too small (need %lld)", (long long)dynpy_pdist_workspace_bytes(n));
        return DYNPY_ERR_WORKSPACE;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_pdist(ux, uy, uz, stride, n, a, b, c, (const i64*)index, dmax, dx, dy, dz, dr, (i64*)atom0,
                             (i64*)atom1, (i64*)projection, (i64*)count, (i64*)ws, st), "pdist_ortho");
    return DYNPY_OK;
}

int dynpy_cart_to_dipolar_f64(const double* dx, const double* dy, const double* dz, const double* dr, int64_t n,
                              double* out, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || (n > 0 && (!dx || !dy || !dz || !dr || !out))) {
        set_error("dynpy_cart_to_dipolar_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_cart_to_dipolar(dx, dy, dz, dr, n, out, d->n_sm, st), "cart_to_dipolar");
    return DYNPY_OK;
}

int dynpy_supercell_f64(const double* coords, int64_t n_atoms, int64_t n_frames, int64_t frame, double a, double* out,
                        dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!coords || !out || n_atoms < 1 || n_frames < 1 || frame < 0 || frame >= n_frames) {
        set_error("dynpy_supercell_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_supercell(coords, n_atoms, n_frames, frame, a, out, st), "supercell");
    return DYNPY_OK;
}

int dynpy_dd_pair_count_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const int32_t* pair_i,
                            const int32_t* pair_j, int64_t n_pairs, double a, double b, double c, double rmax,
                            int64_t* counts, int32_t* frame_hit, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!coords || !pair_i || !pair_j || !counts || n_frames < 1 || n_pairs < 0 || n_atoms < 1) {
        set_error("dynpy_dd_pair_count_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    for (i64 p0 = 0; p0 < n_pairs; p0 += (i64)1 << 30) {
        i64 cnt = n_pairs - p0 < ((i64)1 << 30) ? n_pairs - p0 : ((i64)1 << 30);
        DYNPY_CHECK(launch_dd_pair_count(coords, n_frames, pair_i + p0, pair_j + p0, cnt, a, b, c, rmax,
                                         (i64*)counts + p0, frame_hit, st), "dd pair count");
    }
    return DYNPY_OK;
}

int dynpy_topology_check_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const double* rcov,
                             double bond_extra, const uint8_t* expect, double a, double b, double c, double dmax,
                             uint64_t* out, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!coords || !rcov || !expect || !out || n_atoms < 1 || n_frames < 1 || n_atoms > 65535) {
        set_error("dynpy_topology_check_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    const unsigned long long init[2] = {0ull, ~0ull};
    DYNPY_CHECK(cudaMemcpyAsync(out, init, sizeof(init), cudaMemcpyHostToDevice, st), "topology init");
    DYNPY_CHECK(launch_topology_check(coords, n_atoms, n_frames, 0, n_frames, rcov, bond_extra, expect, a, b, c, dmax,
                                      (unsigned long long*)out, nullptr, 0, st), "topology check");
    return DYNPY_OK;
}

int dynpy_topology_events_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const double* rcov,
                              double bond_extra, const uint8_t* expect, double a, double b, double c, double dmax,
                              int64_t* events, int64_t capacity, uint64_t* out, dynpy_stream_t stream) {
    return dynpy_topology_events_range_f64(coords, n_atoms, n_frames, 0, n_frames, rcov, bond_extra, expect, a, b, c,
                                           dmax, events, capacity, out, stream);
}

int dynpy_topology_events_range_f64(const double* coords, int64_t n_atoms, int64_t n_frames, int64_t frame_begin,
                                    int64_t frame_end, const double* rcov, double bond_extra, const uint8_t* expect,
                                    double a, double b, double c, double dmax, int64_t* events, int64_t capacity,
                                    uint64_t* out, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!coords || !rcov || !expect || !out || !events || capacity < 1 || n_atoms < 1 || n_frames < 1 ||
        n_atoms > 65535 || frame_begin < 0 || frame_end > n_frames || frame_begin > frame_end) {
        set_error("dynpy_topology_events_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    const unsigned long long init[3] = {0ull, ~0ull, 0ull};
    DYNPY_CHECK(cudaMemcpyAsync(out, init, sizeof(init), cudaMemcpyHostToDevice, st), "topology init");
    DYNPY_CHECK(launch_topology_check(coords, n_atoms, n_frames, frame_begin, frame_end, rcov, bond_extra, expect, a, b,
                                      c, dmax, (unsigned long long*)out, (i64*)events, capacity, st), "topology events");
    return DYNPY_OK;
}

static i64 dd_item_bytes(i64 N, i64 M) {
    i64 z = align_up((i64)DD_NSLOTS * 2 * N * 8, 256);
    i64 t = M > DYNPY_ROW_LEN ? (i64)DD_NSLOTS * M * (i64)sizeof(cd) : 0;
    return z + t + 16;
}
int64_t dynpy_dd_workspace_bytes(int64_t n_frames, int64_t M, int64_t items) {
    if (items < 1) items = 1;
    int64_t need = items * dd_item_bytes(n_frames, M) + 1024;
    int64_t fast = dd2_supported(M) ? dd2_workspace_bytes(M, n_frames, items) : 0;
    return need > fast ? need : fast;
}

int dynpy_dd_pairs_to_spectrum_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const int32_t* pair_i,
                                   const int32_t* pair_j, int64_t n_pairs, double a, double b, double c,
                                   double* spectrum, int64_t M, void* ws, int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_M(M, n_frames);
    if (rc) return rc;
    if (!coords || !pair_i || !pair_j || !spectrum || n_pairs < 0 || n_atoms < 1 || !ws) {
        set_error("dynpy_dd_pairs_to_spectrum_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    if (n_pairs == 0) return DYNPY_OK;
    DeviceState* d;
    rc = get_device_state(&d, st);
    if (rc) return rc;
    const i64 N = n_frames;
    {
        const char* sel = getenv("DYNPY_DD_PATH");  // "v1": generic word generator + engine (A/B checks)
        if (dd2_supported(M) && !(sel && sel[0] == 'v' && sel[1] == '1') && ws_bytes >= dd2_workspace_bytes(M, N, 1))
            return dd2_run(coords, N, pair_i, pair_j, n_pairs, a, b, c, spectrum, M, ws, ws_bytes, d->tw4096, d->n_sm, st);
    }
    i64 cap = (ws_bytes - 1024) / dd_item_bytes(N, M);
    if (cap < 1) {
        set_error("dynpy_dd_pairs_to_spectrum_f64: workspace too small (need >= %lld bytes)",
                  (long long)dynpy_dd_workspace_bytes(N, M, 1));
        return DYNPY_ERR_WORKSPACE;
    }
    const i64 items = (n_pairs + 1) / 2;
    if (cap > items) cap = items;
    // carve: bias[2*cap] | Z[cap][11][2][N] | T[cap*11][M]
    char* base = (char*)ws;
    double* bias = (double*)base;
    i64 off = align_up(cap * 16, 256);
    double* Z = (double*)(base + off);
    off += align_up(cap * (i64)DD_NSLOTS * 2 * N * 8, 256);
    cd* T = M > DYNPY_ROW_LEN ? (cd*)(base + off) : nullptr;

    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = spectrum;
    ep.M = M;
    ep.nslots = DD_NSLOTS;
    const int acc_of_slot[DD_NSLOTS] = {0, 1, 1, 2, 2, 3, 4, 5, 5, 6, 6};
    for (int s = 0; s < DD_NSLOTS; ++s) ep.acc_pack |= (unsigned long long)acc_of_slot[s] << (4 * s);

    for (i64 it0 = 0; it0 < items; it0 += cap) {
        const i64 cnt = items - it0 < cap ? items - it0 : cap;
        {
            ProfScope prof(PROF_WORDS, st);
            DYNPY_CHECK(launch_dd_words(coords, n_atoms, N, pair_i, pair_j, 2 * it0, n_pairs, cnt, a, b, c, Z, bias, st),
                        "dd words");
        }
        PlanarLoader ld;
        memset(&ld, 0, sizeof(ld));
        ld.re = Z;
        ld.im = Z + N;
        ld.hi = (i64)DD_NSLOTS * 2 * N;
        ld.lo = 2 * N;
        ld.n_im_valid = (i64)1 << 62;
        ld.nslots = DD_NSLOTS;
        ld.N = N;
        ld.bias = bias;
        ld.bias_slot = 5;
        DYNPY_CHECK(forward_planar_acc(ld, ep, cnt * DD_NSLOTS, M, DD_NSLOTS, T, cnt * DD_NSLOTS, d->tw4096, d->n_sm, st),
                    "dd forward");
    }
    return DYNPY_OK;
}

static i64 dd_ragged_pair_bytes(i64 N, i64 M) {
    i64 w = align_up(14 * N * 8, 256);        // compressed words
    i64 fi = align_up(N * 4, 256);            // rank -> frame
    i64 p = align_up(7 * M * 8, 256) * 2;     // per-series power spectra + their even parts in natural order
    i64 acf = align_up(7 * N * 8, 256);       // per-series ACFs
    i64 small = 7 * 8 * 2 + 16 + 256;         // lens, scale, bias
    return w + fi + p + acf + small;
}
// one spare row of the packed even spectra and of the ACFs (odd pair counts: see dd_pack_rows)
static i64 dd_ragged_extra_bytes(i64 N, i64 M) { return align_up(M * 8, 256) + align_up(N * 8, 256); }

// Transforms the four-step intermediate T of the ragged path holds: 14 for small calls (46 MB at M = 204800: stays in
// L2), 224 from 32 pairs on — with 14 a launch covers 700 rows of 4096 points on 296 CTAs and the path was bound by
// launch tails (66 000 launches of ~20 us per cfg2 step at rmax = 15 bohr).
static i64 dd_ragged_tcap(i64 pairs) { return pairs >= 32 ? 224 : 14; }

int64_t dynpy_dd_ragged_workspace_bytes(int64_t n_frames, int64_t M, int64_t pairs) {
    if (pairs < 1) pairs = 1;
    i64 t = M > DYNPY_ROW_LEN ? dd_ragged_tcap(pairs) * M * (i64)sizeof(cd) : 0;
    return pairs * dd_ragged_pair_bytes(n_frames, M) + t + 4096 + dd_ragged_extra_bytes(n_frames, M);
}

int dynpy_dd_pairs_ragged_to_acf_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const int32_t* pair_i,
                                     const int32_t* pair_j, int64_t n_pairs, double a, double b, double c,
                                     double rmax, double* G, int64_t M, void* ws, int64_t ws_bytes,
                                     dynpy_stream_t stream) {
    return dynpy_dd_pairs_masked_to_acf_f64(coords, n_atoms, n_frames, pair_i, pair_j, n_pairs, a, b, c, rmax, nullptr,
                                            G, nullptr, nullptr, M, ws, ws_bytes, stream);
}

int dynpy_dd_pairs_masked_to_acf_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const int32_t* pair_i,
                                     const int32_t* pair_j, int64_t n_pairs, double a, double b, double c,
                                     double rmax, const uint8_t* mask, double* G, int64_t* pair_frames,
                                     int32_t* frame_hit, int64_t M, void* ws, int64_t ws_bytes,
                                     dynpy_stream_t stream) {
    return dynpy_dd_pairs_ragged_capped_f64(coords, n_atoms, n_frames, pair_i, pair_j, n_pairs, a, b, c, rmax, mask, G,
                                            pair_frames, frame_hit, n_frames, M, ws, ws_bytes, stream);
}

int dynpy_dd_pairs_ragged_capped_f64(const double* coords, int64_t n_atoms, int64_t n_frames, const int32_t* pair_i,
                                     const int32_t* pair_j, int64_t n_pairs, double a, double b, double c,
                                     double rmax, const uint8_t* mask, double* G, int64_t* pair_frames,
                                     int32_t* frame_hit, int64_t n_cap, int64_t M, void* ws, int64_t ws_bytes,
                                     dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cap < 1 || n_cap > n_frames) {
        set_error("dynpy_dd_pairs_ragged_capped_f64: n_cap must lie in [1, n_frames]");
        return DYNPY_ERR_INVALID;
    }
    int rc = check_M(M, n_cap);
    if (rc) return rc;
    if (!coords || !pair_i || !pair_j || !G || n_pairs < 0 || n_atoms < 1 || !ws) {
        set_error("dynpy_dd_pairs_ragged_to_acf_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    if (n_pairs == 0) return DYNPY_OK;
    DeviceState* d;
    rc = get_device_state(&d, st);
    if (rc) return rc;
    const i64 N = n_frames, NC = n_cap;
    i64 T_cap = M > DYNPY_ROW_LEN ? dd_ragged_tcap(n_pairs) : 0;
    i64 tbytes = T_cap * M * (i64)sizeof(cd);
    i64 cap = (ws_bytes - tbytes - 4096 - dd_ragged_extra_bytes(NC, M)) / dd_ragged_pair_bytes(NC, M);
    if (cap < 32 && T_cap > 14) {   // a workspace sized for a small batch: the small T
        T_cap = 14;
        tbytes = T_cap * M * (i64)sizeof(cd);
        cap = (ws_bytes - tbytes - 4096 - dd_ragged_extra_bytes(NC, M)) / dd_ragged_pair_bytes(NC, M);
    }
    if (cap < 1) {
        set_error("dynpy_dd_pairs_ragged_to_acf_f64: workspace too small (need >= %lld bytes)",
                  (long long)dynpy_dd_ragged_workspace_bytes(NC, M, 1));
        return DYNPY_ERR_WORKSPACE;
    }
    if (cap > n_pairs) cap = n_pairs;
    char* base = (char*)ws;
    i64 off = 0;
    auto carve = [&](i64 bytes) { char* p = base + off; off += align_up(bytes, 256); return p; };
    i64* lens = (i64*)carve(cap * 7 * 8);
    double* scale = (double*)carve(cap * 7 * 8);
    double* bias = (double*)carve(cap * 16);
    int* fidx = (int*)carve(cap * NC * 4);
    double* W = (double*)carve(cap * 14 * NC * 8);
    double* P = (double*)carve(cap * 7 * M * 8);
    double* Pn = (double*)carve((cap * 7 + 1) * M * 8);
    double* acf = (double*)carve((cap * 7 + 1) * NC * 8);
    cd* T = tbytes ? (cd*)carve(tbytes) : nullptr;

    for (i64 p0 = 0; p0 < n_pairs; p0 += cap) {
        const i64 cnt = n_pairs - p0 < cap ? n_pairs - p0 : cap;
        DYNPY_CHECK(launch_dd_ragged_words(coords, N, NC, pair_i, pair_j, p0, cnt, a, b, c, rmax, W, fidx, lens, scale,
                                           bias, mask, st), "dd ragged words");
        DYNPY_CHECK(launch_dd_ragged_presence(fidx, lens, NC, cnt, pair_frames ? (i64*)pair_frames + p0 : nullptr,
                                              frame_hit, st), "dd ragged presence");
        PlanarLoader ld;
        memset(&ld, 0, sizeof(ld));
        ld.re = W;
        ld.im = W + NC;
        ld.hi = 14 * NC;
        ld.lo = 2 * NC;
        ld.n_im_valid = (i64)1 << 62;
        ld.nslots = 7;
        ld.N = NC;
        ld.lens = lens;
        ld.bias = bias;
        ld.bias_slot = 3;
        EpiParams ep;
        memset(&ep, 0, sizeof(ep));
        ep.out = P;
        ep.M = M;
        ep.nslots = 7;
        DYNPY_CHECK(forward_planar_power(ld, ep, cnt * 7, M, 7, T, T_cap, d->tw4096, d->n_sm, st), "dd ragged forward");
        // even parts of the 7 spectra of every pair in natural order, two per inverse transform; the rows are
        // permuted (dd_pack_rows) so that the two spectra of a transform have the same magnitude
        const i64 n_rows = dd_pack_rows(cnt);
        DYNPY_CHECK(launch_spectrum_even_natural(P, Pn, cnt * 7, M, st), "dd ragged even parts");
        SpectrumPairLoader sl;
        sl.S = Pn;
        sl.M = M;
        sl.n_spec = n_rows;
        EpiParams e2;
        memset(&e2, 0, sizeof(e2));
        e2.out = acf;
        e2.M = M;
        e2.nslots = 1;
        e2.n_out = NC;
        e2.n_series = n_rows;
        e2.scale = 1.0 / (double)M;
        DYNPY_CHECK(forward_spectrum_real2(sl, e2, n_rows / 2, M, T, T_cap, d->tw4096, d->n_sm, st),
                    "dd ragged inverse");
        DYNPY_CHECK(launch_dd_ragged_scatter(acf, fidx, lens, N, NC, cnt * 7, G, st), "dd ragged scatter");
    }
    return DYNPY_OK;
}

// ------------------------------------------------------------------------------------------ SR
int dynpy_sr_angmom_f64(const double* pos, const double* vel, int64_t n_atoms, int64_t n_frames,
                        const int32_t* mol_atoms, int64_t n_mol, int n_at_mol, const double* mass,
                        const int32_t* axis_atoms, int axis_rule, double a, double b, double c, double inv_dt,
                        double* J, double* omega, double* axes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!pos || !mol_atoms || !mass || !axis_atoms || n_mol < 0 || n_at_mol < 1 || n_frames < (vel ? 1 : 2) ||
        (axis_rule != 0 && axis_rule != 1) || n_atoms < 1) {
        set_error("dynpy_sr_angmom_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_sr_angmom(pos, vel, n_frames, mol_atoms, n_mol, n_at_mol, mass, axis_atoms, axis_rule, a, b, c,
                                 inv_dt, J, omega, axes, st), "sr angmom");
    return DYNPY_OK;
}

// ------------------------------------------------------------------------------------------ Simpson
int dynpy_simpson_f64(const double* y, const double* x, int64_t n, int n_cols, int64_t col_stride, int mode,
                      double* out, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!y || !x || !out || n < 1 || n_cols < 0 || col_stride < n || (mode != 0 && mode != 1)) {
        set_error("dynpy_simpson_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_simpson(y, x, n, n_cols, col_stride, mode, out, st), "simpson");
    return DYNPY_OK;
}

// ------------------------------------------------------------------------------------------ J(omega)
int64_t dynpy_dft_workspace_bytes(int64_t n) { return n >= 1 ? dft_workspace_bytes(n) : 0; }

int dynpy_dft_f64(const double* re, const double* im, int64_t n_series, int64_t n, double* out, int64_t n_out,
                  void* ws, int64_t ws_bytes, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!re || !out || !ws || n_series < 0 || n < 1 || n >= ((int64_t)1 << 30) || n_out < 0 || n_out > n) {
        set_error("dynpy_dft_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    if (n_series == 0 || n_out == 0) return DYNPY_OK;
    return dft_run(re, im, n_series, n, (cd*)out, n_out, ws, ws_bytes, d->n_sm, st);
}

int dynpy_cutoff_bounds_f64(const double* y, int64_t n, int n_cols, int64_t col_stride, int window, double tol,
                            int64_t* bounds, dynpy_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!y || !bounds || n < 1 || n_cols < 0 || col_stride < n || window < 1 || !(tol >= 0.0)) {
        set_error("dynpy_cutoff_bounds_f64: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    DeviceState* d;
    int rc = get_device_state(&d, st);
    if (rc) return rc;
    DYNPY_CHECK(launch_cutoff_bounds(y, n, n_cols, col_stride, window, tol, (i64*)bounds, d->n_sm, st), "cutoff bounds");
    return DYNPY_OK;
}

#pragma GCC visibility pop
}  // extern "C"
