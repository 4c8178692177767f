// kdejs_io.cpp -- multi-threaded ingest of PopPUNK / Pansim distance tables into (n,2) float64.
//
// Replaces, for the path's inputs only: np.loadtxt(file, delimiter='\t') (poppunk_mod.py:360, :264;
// KDE_distance.py:138-139; scripts/distance_to_gridsearch.py:49) and pd.read_csv(file, header=None,
// sep='\t') + "last two columns" (scripts/pairwise_2D_distance.py:38-45; poppunk_mod.py:266-267).
// Numbers are parsed with std::from_chars (correctly rounded, the same values float() gives numpy).
#include "../../include/kdejs.h"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_io_err;

// pandas' default NA tokens (pd.read_csv, the reader scripts/pairwise_2D_distance.py:38-45 and
// poppunk_mod.py:266-267 use): a field spelled like one of these is a missing distance and becomes NaN, which the
// path maps to scaled 0 like the reference (KDE_distance.py:53).  An EMPTY field is missing, too.
bool is_na_token(const char *b, const char *e) {
    static const char *const kNa[] = {"NA", "N/A", "NaN", "nan", "-NaN", "-nan", "NULL", "null", "n/a", "#N/A",
                                      "#NA", "<NA>", "#N/A N/A", "-1.#IND", "1.#IND", "-1.#QNAN", "1.#QNAN", "None"};
    const size_t n = (size_t)(e - b);
    for (const char *tok : kNa)
        if (std::strlen(tok) == n && std::memcmp(tok, b, n) == 0) return true;
    return false;
}

bool parse_double(const char *b, const char *e, double &v) {
    while (b < e && (*b == ' ' || *b == '\r')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\r')) --e;
    if (b >= e || is_na_token(b, e)) { v = std::numeric_limits<double>::quiet_NaN(); return true; }
    if (*b == '+') ++b;                                // from_chars rejects a leading '+'
    if (b >= e) return false;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}

// last two tab-separated fields of [b, e)
bool last_two(const char *b, const char *e, double &x, double &y) {
    const char *p = e;
    const char *t2 = nullptr;          // tab before the last field
    const char *s1 = b;                // start of the field before it
    while (p > b) {
        --p;
        if (*p == '\t') {
            if (!t2) t2 = p;
            else { s1 = p + 1; break; }
        }
    }
    if (!t2) return false;             // fewer than two fields
    return parse_double(s1, t2, x) && parse_double(t2 + 1, e, y);
}
}  // namespace

extern "C" {

const char *kdejs_io_last_error(void) { return g_io_err.c_str(); }

int kdejs_tsv_load(const char *path, int skip_rows, int pinned, int threads, double **out_pts, int64_t *out_rows) {
    if (!path || !out_pts || !out_rows) { g_io_err = "null argument"; return KDEJS_ERR_ARG; }
    FILE *f = std::fopen(path, "rb");
    if (!f) { g_io_err = std::string("cannot open ") + path; return KDEJS_ERR_ARG; }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)std::max(0L, sz) + 1);
    size_t got = sz > 0 ? std::fread(buf.data(), 1, (size_t)sz, f) : 0;
    std::fclose(f);
    buf[got] = '\n';
    const char *base = buf.data(), *end = base + got;
    // line starts (blank lines and '#' comments are skipped, as np.loadtxt does)
    std::vector<const char *> lines;
    lines.reserve(got / 24 + 16);
    for (const char *p = base; p < end;) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p) + 1);
        const char *q = p;
        while (q < nl && (*q == ' ' || *q == '\r' || *q == '\t')) ++q;
        if (q < nl && *q != '#') lines.push_back(p);
        p = nl + 1;
    }
    size_t first = 0;
    if (skip_rows < 0) {                 // auto: drop one header line if its last two fields are not numbers
        if (!lines.empty()) {
            const char *nl = (const char *)std::memchr(lines[0], '\n', (size_t)(end - lines[0]) + 1);
            double x, y;
            // a header is a first line whose last two fields are neither numbers nor missing-value tokens
            if (!last_two(lines[0], nl, x, y)) first = 1;
        }
    } else first = std::min((size_t)skip_rows, lines.size());
    const int64_t n = (int64_t)(lines.size() - first);
    if (n <= 0) { g_io_err = std::string("no data rows in ") + path; return KDEJS_ERR_ARG; }
    double *out = nullptr;
    if (pinned) out = (double *)kdejs_host_alloc((size_t)n * 16);
    else out = (double *)std::malloc((size_t)n * 16);
    if (!out) { g_io_err = "allocation failed"; return KDEJS_ERR_CUDA; }
    int nt = threads > 0 ? threads : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    nt = (int)std::min<int64_t>(nt, std::max<int64_t>(1, n / 4096));
    std::vector<int64_t> bad(nt, -1);
    auto work = [&](int t) {
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        for (int64_t i = lo; i < hi; ++i) {
            const char *b = lines[first + (size_t)i];
            const char *nl = (const char *)std::memchr(b, '\n', (size_t)(end - b) + 1);
            double x, y;
            if (!last_two(b, nl, x, y)) { if (bad[t] < 0) bad[t] = i; x = y = 0.0; }
            out[2 * i] = x;
            out[2 * i + 1] = y;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < nt; ++t)
        if (bad[t] >= 0) {
            g_io_err = std::string(path) + ": could not parse two numeric columns in data row " + std::to_string(bad[t] + 1);
            kdejs_tsv_free(out, pinned);
            return KDEJS_ERR_ARG;
        }
    *out_pts = out;
    *out_rows = n;
    return KDEJS_OK;
}

void kdejs_tsv_free(double *pts, int pinned) {
    if (!pts) return;
    if (pinned) kdejs_host_free(pts);
    else std::free(pts);
}

}  // extern "C"
