// kdejs_io.cpp -- multi-threaded ingest of PopPUNK / Pansim distance tables into (n,2) float64.
//
// Replaces, for the path's inputs only: np.loadtxt(file, delimiter='\t') (poppunk_mod.py:360, :264;
// KDE_distance.py:138-139; scripts/distance_to_gridsearch.py:49) and pd.read_csv(file, header=None,
// sep='\t') + "last two columns" (scripts/pairwise_2D_distance.py:38-45; poppunk_mod.py:266-267).
// Numbers are parsed with std::from_chars (correctly rounded, the same values float() gives numpy).
//
// Layout of the work: the file is read whole; it is cut into one byte range per thread, each range starting at a
// line start; pass 1 counts the data rows of every range (blank lines and '#' comments do not count), which fixes
// where each range's rows land in the output; pass 2 parses every range straight into its place.  A field goes to
// from_chars first and is compared with pandas' missing-value spellings only when that fails -- the first version
// checked the 18 spellings before every number and spent four fifths of its time there (100 000 rows of the
// reference's example table: 21.7 -> 6.8 ms on one core).  from_chars itself (Eisel-Lemire, ~25 ns per number) is most of
// what is left; a hand-written conversion (one IEEE operation for up to 15 digits, one x87 80-bit operation with a
// guard band for 16-19) was measured SLOWER than it (8.0 vs 6.8 ms) and removed.
#include "../../include/kdejs.h"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
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

inline bool parse_double(const char *b, const char *e, double &v) {
    while (b < e && (*b == ' ' || *b == '\r')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\r')) --e;
    if (b >= e) { v = std::numeric_limits<double>::quiet_NaN(); return true; }      // empty field: missing
    const char *s = (*b == '+') ? b + 1 : b;           // from_chars rejects a leading '+'
    if (s < e) {
        auto r = std::from_chars(s, e, v);
        if (r.ec == std::errc() && r.ptr == e) return true;
        if (r.ec == std::errc::result_out_of_range && r.ptr == e) {      // 1e400 -> inf, 1e-400 -> 0 like float()
            const std::string z(s, e);
            char *stop = nullptr;
            v = std::strtod(z.c_str(), &stop);
            return stop == z.c_str() + z.size();
        }
    }
    if (is_na_token(b, e)) { v = std::numeric_limits<double>::quiet_NaN(); return true; }
    return false;
}

// last two tab-separated fields of [b, e)
inline bool last_two(const char *b, const char *e, double &x, double &y) {
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

// a data line: not blank, not a '#' comment (np.loadtxt skips both)
inline bool is_data_line(const char *p, const char *nl) {
    while (p < nl && (*p == ' ' || *p == '\r' || *p == '\t')) ++p;
    return p < nl && *p != '#';
}

struct Table {                 // the file in memory (one '\n' appended) and where its data rows begin
    std::unique_ptr<char[]> buf;
    const char *data = nullptr, *end = nullptr;
};

int read_table(const char *path, int skip_rows, Table &t) {
    FILE *f = std::fopen(path, "rb");
    if (!f) { g_io_err = std::string("cannot open ") + path; return KDEJS_ERR_ARG; }
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    t.buf.reset(new char[(size_t)std::max(0L, sz) + 1]);
    const size_t got = sz > 0 ? std::fread(t.buf.get(), 1, (size_t)sz, f) : 0;
    std::fclose(f);
    t.buf[got] = '\n';
    const char *p = t.buf.get();
    t.end = p + got;
    // rows dropped from the top: `skip_rows` data lines, or (auto, -1) one line if its last two fields are neither
    // numbers nor missing-value tokens (a header)
    int to_skip = std::max(skip_rows, 0);
    bool probe_header = skip_rows < 0;
    while (p < t.end && (to_skip > 0 || probe_header)) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(t.end - p) + 1);
        if (is_data_line(p, nl)) {
            if (probe_header) {
                double x, y;
                if (last_two(p, nl, x, y)) break;          // numeric first row: nothing to drop
                probe_header = false;
            } else --to_skip;
        }
        p = nl + 1;
    }
    t.data = std::min(p, t.end);
    return KDEJS_OK;
}

// Cuts [data, end) into nt ranges that start at line starts, counts each range's data rows (pass 1).
void cut_and_count(const Table &t, int nt, std::vector<const char *> &cut, std::vector<int64_t> &rows) {
    cut.assign((size_t)nt + 1, t.end);
    cut[0] = t.data;
    const size_t bytes = (size_t)(t.end - t.data);
    for (int k = 1; k < nt; ++k) {
        const char *p = t.data + bytes * (size_t)k / (size_t)nt;
        p = std::max(p, cut[k - 1]);
        if (p < t.end && p > t.data && p[-1] != '\n') {
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(t.end - p) + 1);
            p = std::min(nl + 1, t.end);
        }
        cut[k] = p;
    }
    rows.assign((size_t)nt, 0);
    auto count = [&](int k) {
        int64_t n = 0;
        for (const char *p = cut[k]; p < cut[k + 1];) {
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(t.end - p) + 1);
            n += is_data_line(p, nl) ? 1 : 0;
            p = nl + 1;
        }
        rows[k] = n;
    };
    if (nt == 1) count(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(count, k);
        for (auto &x : th) x.join();
    }
}

// Pass 2: every range parses its rows into out + 2 * (rows of the ranges before it).  Returns the first bad data
// row (0-based) or -1.
int64_t parse_ranges(const Table &t, int nt, const std::vector<const char *> &cut, const std::vector<int64_t> &rows,
                     double *out) {
    std::vector<int64_t> first(nt + 1, 0), bad(nt, -1);
    for (int k = 0; k < nt; ++k) first[k + 1] = first[k] + rows[k];
    auto work = [&](int k) {
        int64_t i = first[k];
        for (const char *p = cut[k]; p < cut[k + 1];) {
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(t.end - p) + 1);
            if (is_data_line(p, nl)) {
                double x, y;
                if (!last_two(p, nl, x, y)) { if (bad[k] < 0) bad[k] = i; x = y = 0.0; }
                out[2 * i] = x;
                out[2 * i + 1] = y;
                ++i;
            }
            p = nl + 1;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(work, k);
        for (auto &x : th) x.join();
    }
    for (int k = 0; k < nt; ++k)
        if (bad[k] >= 0) return bad[k];
    return -1;
}

int thread_count(int threads, size_t bytes) {
    int nt = threads > 0 ? threads : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)nt, bytes / (128 * 1024)));     // a thread per >= 128 KB
}
}  // namespace

extern "C" {

const char *kdejs_io_last_error(void) { return g_io_err.c_str(); }

int kdejs_tsv_load_into(const char *path, int skip_rows, int threads, double *dst, int64_t capacity_rows,
                        int64_t *out_rows) {
    if (!path || !out_rows || (capacity_rows > 0 && !dst)) { g_io_err = "null argument"; return KDEJS_ERR_ARG; }
    *out_rows = 0;
    Table t;
    if (int rc = read_table(path, skip_rows, t)) return rc;
    const int nt = thread_count(threads, (size_t)(t.end - t.data));
    std::vector<const char *> cut;
    std::vector<int64_t> rows;
    cut_and_count(t, nt, cut, rows);
    int64_t n = 0;
    for (int64_t r : rows) n += r;
    *out_rows = n;
    if (n <= 0) { g_io_err = std::string("no data rows in ") + path; return KDEJS_ERR_ARG; }
    if (n > capacity_rows) {
        g_io_err = std::string(path) + ": " + std::to_string(n) + " data rows do not fit the buffer (" +
                   std::to_string(capacity_rows) + " rows)";
        return KDEJS_ERR_CAPACITY;
    }
    const int64_t bad = parse_ranges(t, nt, cut, rows, dst);
    if (bad >= 0) {
        g_io_err = std::string(path) + ": could not parse two numeric columns in data row " + std::to_string(bad + 1);
        return KDEJS_ERR_ARG;
    }
    return KDEJS_OK;
}

int kdejs_tsv_load(const char *path, int skip_rows, int pinned, int threads, double **out_pts, int64_t *out_rows) {
    if (!path || !out_pts || !out_rows) { g_io_err = "null argument"; return KDEJS_ERR_ARG; }
    Table t;
    if (int rc = read_table(path, skip_rows, t)) return rc;
    const int nt = thread_count(threads, (size_t)(t.end - t.data));
    std::vector<const char *> cut;
    std::vector<int64_t> rows;
    cut_and_count(t, nt, cut, rows);
    int64_t n = 0;
    for (int64_t r : rows) n += r;
    if (n <= 0) { g_io_err = std::string("no data rows in ") + path; return KDEJS_ERR_ARG; }
    double *out = nullptr;
    if (pinned) out = (double *)kdejs_host_alloc((size_t)n * 16);
    else out = (double *)std::malloc((size_t)n * 16);
    if (!out) { g_io_err = "allocation failed"; return KDEJS_ERR_CUDA; }
    const int64_t bad = parse_ranges(t, nt, cut, rows, out);
    if (bad >= 0) {
        g_io_err = std::string(path) + ": could not parse two numeric columns in data row " + std::to_string(bad + 1);
        kdejs_tsv_free(out, pinned);
        return KDEJS_ERR_ARG;
    }
    *out_pts = out;
    *out_rows = n;
    return KDEJS_OK;
}

/* Where the data rows of a table begin, for text that is already in memory (the device parser gets this offset):
 * the same rule as kdejs_tsv_load -- `skip_rows` data lines are dropped, or (-1) one line if its last two fields are
 * not numeric.  The text need not end with a newline. */
int64_t kdejs_tsv_data_offset(const char *text, int64_t bytes, int skip_rows) {
    if (!text || bytes <= 0) return 0;
    const char *p = text, *end = text + bytes;
    int to_skip = std::max(skip_rows, 0);
    bool probe_header = skip_rows < 0;
    while (p < end && (to_skip > 0 || probe_header)) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        if (!nl) nl = end;
        if (is_data_line(p, nl)) {
            if (probe_header) {
                double x, y;
                if (last_two(p, nl, x, y)) break;
                probe_header = false;
            } else --to_skip;
        }
        p = nl < end ? nl + 1 : end;
    }
    return (int64_t)(p - text);
}

/* Reads n files into slots of `slot_bytes` bytes each (dst + i * slot_bytes), `threads` files at a time: plain
 * read(2) into the caller's (page-locked) arena, no parsing.  out_bytes[i] = bytes read, or -(size of the file) when
 * it does not fit its slot (nothing is read then), or -1 when the file cannot be opened. */
int kdejs_files_read(const char *const *paths, int n, char *dst, int64_t slot_bytes, int64_t *out_bytes, int threads) {
    if (n < 0 || (n > 0 && (!paths || !dst || !out_bytes)) || slot_bytes <= 0) { g_io_err = "bad argument"; return KDEJS_ERR_ARG; }
    const int nt = std::max(1, std::min(threads > 0 ? threads : 8, n));
    auto work = [&](int t) {
        for (int i = t; i < n; i += nt) {
            out_bytes[i] = -1;
            FILE *f = std::fopen(paths[i], "rb");
            if (!f) continue;
            std::setvbuf(f, nullptr, _IONBF, 0);
            std::fseek(f, 0, SEEK_END);
            const long sz = std::ftell(f);
            std::fseek(f, 0, SEEK_SET);
            if (sz > slot_bytes) out_bytes[i] = -(int64_t)sz;
            else out_bytes[i] = sz > 0 ? (int64_t)std::fread(dst + (size_t)i * (size_t)slot_bytes, 1, (size_t)sz, f) : 0;
            std::fclose(f);
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    return KDEJS_OK;
}

/* Every number of a small text file -- the simulator's gene-frequency vector (poppunk_mod.py:361: np.loadtxt of
 * "<outpref>_freqs.txt", one value per line or per tab) -- as one malloc'ed vector: fields separated by blanks, tabs
 * or newlines, '#' starts a comment that runs to the end of its line.  Free with kdejs_tsv_free(ptr, 0). */
int kdejs_numbers_load(const char *path, double **out, int64_t *out_n) {
    if (!path || !out || !out_n) { g_io_err = "null argument"; return KDEJS_ERR_ARG; }
    *out = nullptr;
    *out_n = 0;
    Table t;
    if (int rc = read_table(path, 0, t)) return rc;
    const char *p = t.buf.get(), *end = t.end;
    std::vector<double> v;
    v.reserve((size_t)(end - p) / 8 + 16);
    while (p < end) {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
        if (p >= end) break;
        if (*p == '#') { while (p < end && *p != '\n') ++p; continue; }
        const char *q = p;
        while (q < end && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r' && *q != '#') ++q;
        double x;
        if (!parse_double(p, q, x)) {
            g_io_err = std::string(path) + ": could not parse number " + std::to_string(v.size() + 1) + " ('" + std::string(p, q) + "')";
            return KDEJS_ERR_ARG;
        }
        v.push_back(x);
        p = q;
    }
    if (v.empty()) { g_io_err = std::string("no numbers in ") + path; return KDEJS_ERR_ARG; }
    double *mem = (double *)std::malloc(v.size() * sizeof(double));
    if (!mem) { g_io_err = "allocation failed"; return KDEJS_ERR_CUDA; }
    std::memcpy(mem, v.data(), v.size() * sizeof(double));
    *out = mem;
    *out_n = (int64_t)v.size();
    return KDEJS_OK;
}

void kdejs_tsv_free(double *pts, int pinned) {
    if (!pts) return;
    if (pinned) kdejs_host_free(pts);
    else std::free(pts);
}

}  // extern "C"
