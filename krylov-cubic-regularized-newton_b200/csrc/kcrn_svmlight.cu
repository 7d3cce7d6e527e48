// LIBSVM / svmlight text -> CSR on the host.  The reference driver reads its datasets
// with sklearn.datasets.load_svmlight_file (cubic_newton.py:52); this is the same grammar -- "label [qid:n]
// idx:value ... [# comment]", blank lines skipped, stored zeros kept, indices as written (the zero/one-based
// decision and the sort are the caller's, as in sklearn) -- parsed by all host threads: the text is cut at line
// boundaries, every thread counts and then fills its own slice of the output arrays.  Host code only; no device.
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "kcrn.h"
#include "kcrn_internal.cuh"

namespace kcrn {
namespace {

struct Slice {
    const char *lo, *hi;
    int64_t rows = 0, nnz = 0;
    int64_t bad_line = -1;       // line (within the slice) that failed to parse
    bool bad_order = false;      // ... because its indices are not strictly increasing (sklearn rejects those too)
    int32_t min_idx = INT32_MAX, max_idx = -1;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// end of the payload of the line starting at p: the newline, the end of the text or a '#'
inline const char *payload_end(const char *p, const char *line_end) {
    const char *h = (const char *)memchr(p, '#', (size_t)(line_end - p));
    return h ? h : line_end;
}

std::vector<Slice> cut(const char *text, int64_t len, int threads) {
    std::vector<Slice> s;
    const char *end = text + len, *p = text;
    for (int t = 0; t < threads && p < end; ++t) {
        const char *want = text + (len * (int64_t)(t + 1)) / threads;
        if (want < p) want = p;
        const char *q = t + 1 == threads ? end : (const char *)memchr(want, '\n', (size_t)(end - want));
        q = q ? (q < end ? q + 1 : end) : end;
        Slice sl;
        sl.lo = p; sl.hi = q;
        s.push_back(sl);
        p = q;
    }
    return s;
}

// One pass over a slice.  fill = false: count rows and stored entries.  fill = true: write them.
void walk(Slice &sl, bool fill, double *labels, int64_t *indptr, int32_t *indices, double *data, int64_t row0,
          int64_t nz0) {
    const char *p = sl.lo;
    int64_t row = row0, nz = nz0, line = 0;
    while (p < sl.hi) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(sl.hi - p));
        const char *line_end = nl ? nl : sl.hi;
        const char *e = payload_end(p, line_end);
        const char *c = p;
        while (c < e && is_space(*c)) ++c;
        if (c < e) {
            // label
            char *after = nullptr;
            errno = 0;
            const double y = strtod(c, &after);
            if (after == c || after > e || (after < e && !is_space(*after))) { sl.bad_line = line; return; }
            if (fill) labels[row] = y;
            c = after;
            bool first_token = true;
            long long prev = -1;
            for (;;) {
                while (c < e && is_space(*c)) ++c;
                if (c >= e) break;
                if (first_token && e - c > 4 && memcmp(c, "qid:", 4) == 0) {     // ranking files: skipped here
                    while (c < e && !is_space(*c)) ++c;
                    first_token = false;
                    continue;
                }
                first_token = false;
                char *colon = nullptr;
                const long long idx = strtoll(c, &colon, 10);
                if (colon == c || colon >= e || *colon != ':' || idx < 0 || idx > INT32_MAX) { sl.bad_line = line; return; }
                if (idx <= prev) { sl.bad_line = line; sl.bad_order = true; return; }
                prev = idx;
                const char *vs = colon + 1;
                const double v = strtod(vs, &after);
                if (after == vs || after > e || (after < e && !is_space(*after))) { sl.bad_line = line; return; }
                if (fill) { indices[nz] = (int32_t)idx; data[nz] = v; }
                if ((int32_t)idx < sl.min_idx) sl.min_idx = (int32_t)idx;
                if ((int32_t)idx > sl.max_idx) sl.max_idx = (int32_t)idx;
                ++nz;
                c = after;
            }
            ++row;
            if (fill) indptr[row] = nz;
        }
        ++line;
        p = nl ? nl + 1 : sl.hi;
    }
    sl.rows = row - row0;
    sl.nnz = nz - nz0;
}

int run(const char *text, int64_t len, int threads, bool fill, int64_t n_rows, int64_t nnz, double *labels,
        int64_t *indptr, int32_t *indices, double *data, int64_t *rows_out, int64_t *nnz_out, int32_t *min_out,
        int32_t *max_out) {
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (len < (int64_t)threads * 4096) threads = 1;
    std::vector<Slice> sl = cut(text, len, threads);
    auto pass = [&](bool f, const std::vector<int64_t> &r0, const std::vector<int64_t> &z0) {
        std::vector<std::thread> pool;
        for (size_t t = 1; t < sl.size(); ++t)
            pool.emplace_back([&, t]() { walk(sl[t], f, labels, indptr, indices, data, r0[t], z0[t]); });
        if (!sl.empty()) walk(sl[0], f, labels, indptr, indices, data, r0[0], z0[0]);
        for (auto &th : pool) th.join();
    };
    std::vector<int64_t> r0(sl.size() + 1, 0), z0(sl.size() + 1, 0);
    pass(false, r0, z0);
    int64_t lines_before = 0;
    for (size_t t = 0; t < sl.size(); ++t) {
        if (sl[t].bad_line >= 0) {
            for (size_t u = 0; u < t; ++u)
                for (const char *q = sl[u].lo; q < sl[u].hi; ++q) lines_before += *q == '\n';
            set_error(sl[t].bad_order ? "svmlight: feature indices should be sorted and unique (line %lld)"
                                      : "svmlight: cannot parse line %lld",
                      (long long)(lines_before + sl[t].bad_line + 1));
            return KCRN_E_ARG;
        }
        r0[t + 1] = r0[t] + sl[t].rows;
        z0[t + 1] = z0[t] + sl[t].nnz;
    }
    const int64_t rows = r0[sl.size()], nz = z0[sl.size()];
    if (rows_out) *rows_out = rows;
    if (nnz_out) *nnz_out = nz;
    int32_t mn = INT32_MAX, mx = -1;
    for (auto &s : sl) { if (s.min_idx < mn) mn = s.min_idx; if (s.max_idx > mx) mx = s.max_idx; }
    if (min_out) *min_out = nz ? mn : 0;
    if (max_out) *max_out = mx;
    if (!fill) return KCRN_OK;
    if (rows != n_rows || nz != nnz) {
        set_error("svmlight: arrays sized for %lld rows / %lld entries, the text has %lld / %lld", (long long)n_rows,
                  (long long)nnz, (long long)rows, (long long)nz);
        return KCRN_E_ARG;
    }
    indptr[0] = 0;
    pass(true, r0, z0);
    return KCRN_OK;
}

}  // namespace
}  // namespace kcrn

extern "C" {

int kcrn_svmlight_count(const char *text, int64_t len, int threads, int64_t *n_rows, int64_t *nnz,
                        int32_t *min_index, int32_t *max_index) {
    if (!text || len < 0 || !n_rows || !nnz) { kcrn::set_error("kcrn_svmlight_count: null argument"); return KCRN_E_ARG; }
    return kcrn::run(text, len, threads, false, 0, 0, nullptr, nullptr, nullptr, nullptr, n_rows, nnz, min_index, max_index);
}

int kcrn_svmlight_parse(const char *text, int64_t len, int threads, int64_t n_rows, int64_t nnz, double *labels,
                        int64_t *indptr, int32_t *indices, double *data) {
    if (!text || len < 0 || !labels || !indptr || (nnz > 0 && (!indices || !data))) {
        kcrn::set_error("kcrn_svmlight_parse: null argument");
        return KCRN_E_ARG;
    }
    return kcrn::run(text, len, threads, true, n_rows, nnz, labels, indptr, indices, data, nullptr, nullptr, nullptr, nullptr);
}

}  // extern "C"
