// dynpy_b200 — native trajectory ingestion (host side): text blocks -> fp64 [frame][atom][3].
//
// Replaces the text handling of parseMD.parse_xyz / parse_tinker_md (reference parseMD.py:177-278:
// pd.read_csv with a per-line skiprows lambda, then d_to_e and /0.529177 per element): the file is
// memory-mapped, line starts of the selected frames are found by a parallel newline scan, and the
// frames are parsed by a pool of threads straight into the array the GPU copy reads.  Numbers are
// converted with std::from_chars (correctly rounded, same value as Python's float()), Fortran 'D'
// exponents are accepted (d_to_e, parseMD.py:325-329), values are DIVIDED by `scale_div` exactly as
// the reference does, so the result is bit-identical to the reference's arrays.
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dynpy_b200.h"

namespace dynpy {
void set_error(const char* fmt, ...);
}

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

inline const char* skip_ws(const char* s, const char* e) {
    while (s < e && (*s == ' ' || *s == '\t' || *s == '\r')) ++s;
    return s;
}
inline const char* skip_tok(const char* s, const char* e) {
    while (s < e && *s != ' ' && *s != '\t' && *s != '\r' && *s != '\n') ++s;
    return s;
}

// one floating-point token [s, t): from_chars, with a D/d -> E rewrite when needed
inline bool parse_double(const char* s, const char* t, double& v) {
    if (t - s > 0 && *s == '+') ++s;  // from_chars rejects a leading '+', float() accepts it
    auto r = std::from_chars(s, t, v);
    if (r.ec == std::errc() && r.ptr == t) return true;
    char buf[64];
    const size_t len = (size_t)(t - s);
    if (len == 0 || len >= sizeof(buf)) return false;
    for (size_t i = 0; i < len; ++i) buf[i] = (s[i] == 'D' || s[i] == 'd') ? 'E' : s[i];
    r = std::from_chars(buf, buf + len, v);
    return r.ec == std::errc() && r.ptr == buf + len;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

static int parse_blocks(const char* path, int64_t nat, int32_t header_lines, int32_t first_col,
                        const int64_t* printed_host, int64_t n_printed, double scale_div, double* out_host,
                        char* symbols_host, double* header_first_host, int32_t n_threads);

int dynpy_parse_md_blocks_host(const char* path, int64_t nat, int32_t header_lines, int32_t first_col,
                               const int64_t* printed_host, int64_t n_printed, double scale_div, double* out_host,
                               char* symbols_host, int32_t n_threads) {
    return parse_blocks(path, nat, header_lines, first_col, printed_host, n_printed, scale_div, out_host, symbols_host,
                        nullptr, n_threads);
}

int dynpy_parse_md_blocks_hdr_host(const char* path, int64_t nat, int32_t header_lines, int32_t first_col,
                                   const int64_t* printed_host, int64_t n_printed, double scale_div, double* out_host,
                                   char* symbols_host, double* header_first_host, int32_t n_threads) {
    return parse_blocks(path, nat, header_lines, first_col, printed_host, n_printed, scale_div, out_host, symbols_host,
                        header_first_host, n_threads);
}

// first_col = index of the symbol token (the coordinates are the three tokens after it), or -1 for lines that are
// "x y z" only (QE .pos); header_first_host (optional, needs header_lines >= 1): first token of each selected
// block's first header line as a number (the step id of a QE .pos / .vel header)
static int parse_blocks(const char* path, int64_t nat, int32_t header_lines, int32_t first_col,
                        const int64_t* printed_host, int64_t n_printed, double scale_div, double* out_host,
                        char* symbols_host, double* header_first_host, int32_t n_threads) {
    if (!path || nat < 1 || header_lines < 0 || first_col < -1 || !printed_host || n_printed < 1 || !out_host ||
        !(scale_div > 0.0) || (header_first_host && header_lines < 1)) {
        dynpy::set_error("dynpy_parse_md_blocks_host: invalid arguments");
        return DYNPY_ERR_INVALID;
    }
    for (int64_t k = 0; k < n_printed; ++k) {
        if (printed_host[k] < 1 || (k > 0 && printed_host[k] <= printed_host[k - 1])) {
            dynpy::set_error("dynpy_parse_md_blocks_host: printed frames must be 1-indexed and ascending");
            return DYNPY_ERR_INVALID;
        }
    }
    Mapped m;
    m.fd = open(path, O_RDONLY);
    if (m.fd < 0) {
        dynpy::set_error("cannot open %s", path);
        return DYNPY_ERR_INVALID;
    }
    struct stat st;
    if (fstat(m.fd, &st) != 0 || st.st_size <= 0) {
        dynpy::set_error("%s is empty or unreadable", path);
        return DYNPY_ERR_INVALID;
    }
    m.n = (size_t)st.st_size;
    void* mp = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
    if (mp == MAP_FAILED) {
        m.n = 0;
        dynpy::set_error("cannot map %s", path);
        return DYNPY_ERR_INVALID;
    }
    m.p = (const char*)mp;
    madvise(mp, m.n, MADV_SEQUENTIAL);
    const char* base = m.p;
    const char* end = m.p + m.n;

    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    const int64_t blk = nat + header_lines;
    const int64_t last_line = printed_host[n_printed - 1] * blk;  // lines needed (exclusive)

    // ---- pass 1: newline counts per chunk (parallel), then the byte offset of every selected block
    const size_t nchunk = (size_t)nt * 4;
    const size_t csz = (m.n + nchunk - 1) / nchunk;
    std::vector<int64_t> cnt(nchunk + 1, 0);
    {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t c; (c = next.fetch_add(1)) < nchunk;) {
                const char* s = base + c * csz;
                const char* e = s + csz < end ? s + csz : end;
                int64_t n = 0;
                while (s < e) {
                    const char* q = (const char*)memchr(s, '\n', (size_t)(e - s));
                    if (!q) break;
                    ++n;
                    s = q + 1;
                }
                cnt[c + 1] = n;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    for (size_t c = 0; c < nchunk; ++c) cnt[c + 1] += cnt[c];  // newlines before chunk c+1
    const int64_t total_lines = cnt[nchunk] + (end[-1] == '\n' ? 0 : 1);
    if (total_lines < last_line) {
        dynpy::set_error("%s holds fewer than %lld printed frames of %lld atoms", path, (long long)printed_host[n_printed - 1],
                         (long long)nat);
        return DYNPY_ERR_INVALID;
    }
    // start offset of line L (0-indexed): scan inside the chunk that contains newline number L
    auto line_start = [&](int64_t L) -> const char* {
        if (L == 0) return base;
        size_t lo = 0, hi = nchunk;  // largest c with cnt[c] < L  -> newline number L (1-indexed) lies in chunk c
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) / 2;
            if (cnt[mid] < L) lo = mid; else hi = mid;
        }
        const char* s = base + lo * csz;
        const char* e = s + csz < end ? s + csz : end;
        int64_t need = L - cnt[lo];
        while (s < e) {
            const char* q = (const char*)memchr(s, '\n', (size_t)(e - s));
            if (!q) break;
            s = q + 1;
            if (--need == 0) return s;
        }
        return nullptr;
    };

    // ---- pass 2: parse the selected frames.  The frames are cut into contiguous pieces; a piece locates its first
    // block through the newline index once and then walks forward (frames not selected in between are skipped line
    // by line), so the text is scanned once — locating every frame through the index re-scanned half a chunk
    // (tens of MB) per frame.
    const int64_t npieces = n_printed < (int64_t)nt * 8 ? n_printed : (int64_t)nt * 8;
    std::atomic<int64_t> nextp{0};
    std::atomic<int> bad{0};
    std::string errmsg;
    std::atomic<bool> err_set{false};
    auto fail = [&](const std::string& msg) {
        bad.store(1);
        bool exp = false;
        if (err_set.compare_exchange_strong(exp, true)) errmsg = msg;
    };
    auto skip_lines = [&](const char* s, int64_t n) -> const char* {
        while (n > 0 && s < end) {
            const char* q = (const char*)memchr(s, '\n', (size_t)(end - s));
            if (!q) return n == 1 ? end : nullptr;
            s = q + 1;
            --n;
        }
        return n == 0 ? s : nullptr;
    };
    auto parse = [&]() {
      for (int64_t piece; (piece = nextp.fetch_add(1)) < npieces && !bad.load();) {
        const int64_t k0 = n_printed * piece / npieces, k1 = n_printed * (piece + 1) / npieces;
        const char* s = nullptr;
        for (int64_t k = k0; k < k1 && !bad.load(); ++k) {
            s = (k == k0) ? line_start((printed_host[k] - 1) * blk)
                          : skip_lines(s, (printed_host[k] - printed_host[k - 1] - 1) * blk);
            if (!s) { fail("internal: line index out of range"); return; }
            if (header_first_host) {
                const char* le = (const char*)memchr(s, '\n', (size_t)(end - s));
                if (!le) le = end;
                const char* c = skip_ws(s, le);
                const char* t = skip_tok(c, le);
                double v;
                if (t == c || !parse_double(c, t, v)) {
                    fail("frame " + std::to_string(printed_host[k]) + ": cannot parse the header token '" +
                         std::string(c, (size_t)(t - c)) + "'");
                    return;
                }
                header_first_host[k] = v;
            }
            for (int h = 0; h < header_lines; ++h) {
                const char* le = (const char*)memchr(s, '\n', (size_t)(end - s));
                s = le ? le + 1 : end;
            }
            double* out = out_host + k * nat * 3;
            for (int64_t a = 0; a < nat; ++a) {
                const char* le = (const char*)memchr(s, '\n', (size_t)(end - s));
                if (!le) le = end;
                const char* c = skip_ws(s, le);
                for (int col = 0; col < first_col; ++col) c = skip_ws(skip_tok(c, le), le);
                const char* t = first_col < 0 ? c : skip_tok(c, le);   // no symbol token: the numbers start here
                if (first_col >= 0 && t == c) { fail("frame " + std::to_string(printed_host[k]) + ": short line for atom " + std::to_string(a)); return; }
                if (k == 0 && symbols_host && first_col >= 0) {
                    size_t len = (size_t)(t - c) < 7 ? (size_t)(t - c) : 7;
                    memset(symbols_host + a * 8, 0, 8);
                    memcpy(symbols_host + a * 8, c, len);
                }
                for (int d = 0; d < 3; ++d) {
                    c = skip_ws(t, le);
                    t = skip_tok(c, le);
                    double v;
                    if (t == c || !parse_double(c, t, v)) {
                        fail("frame " + std::to_string(printed_host[k]) + ", atom " + std::to_string(a) + ": cannot parse '" +
                             std::string(c, (size_t)(t - c)) + "'");
                        return;
                    }
                    out[a * 3 + d] = v / scale_div;
                }
                s = le < end ? le + 1 : end;
            }
        }
      }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(parse);
        parse();
        for (auto& t : th) t.join();
    }
    if (bad.load()) {
        dynpy::set_error("%s: %s", path, errmsg.c_str());
        return DYNPY_ERR_INVALID;
    }
    return DYNPY_OK;
}

#pragma GCC visibility pop
}  // extern "C"
