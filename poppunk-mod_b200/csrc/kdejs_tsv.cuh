// kdejs_tsv.cuh -- distance tables parsed ON THE DEVICE: the text of a wave of files goes over PCIe as it is on disk
// (2.8 MB per 100 000-row table instead of 1.6 MB of doubles -- the link has room; the host's cores do not: one core
// parses 150 such tables per second, and a sweep from disk was bound by that, not by the GPU).
//
// Replaces, for sweeps: np.loadtxt(file, delimiter='\t') per file (scripts/distance_to_gridsearch.py:49,
// poppunk_mod.py:360) and pandas' reader + last two columns (scripts/pairwise_2D_distance.py:38-45).
// Same rules as the host reader (kdejs_io.cpp): the LAST TWO tab-separated fields of every data line; blank lines and
// '#' comments are skipped; blanks and '\r' around a field are ignored; an EMPTY field is NaN.  Numbers are converted
// by kdejs_decimal.h (integer arithmetic, correctly rounded: the bits float() gives).  Anything that header does not
// cover -- missing-value spellings, inf / nan, more than 19 digits, large exponents -- flags the FILE, and the caller
// sends that file through the host reader instead: no result ever comes from an approximate conversion.
//
// Three launches per wave of files, grid (chunks, files), 256 threads x 32 bytes = 8 KB of text per block:
//   k_tsv_count   data lines starting in each chunk
//   k_tsv_scan    exclusive scan of a file's chunk counts -> first row of each chunk, rows of the file
//   k_tsv_parse   the chunk is staged in shared memory (16-byte loads); every line start parses its line from there and
//                 writes row (x, y) to its place; the table's column bounds are folded with integer atomics
// HBM-bound byte work: the text is read twice (count, parse) with 16-byte loads, the rows are written once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kdejs_decimal.h"

namespace kdejs {

constexpr int kTsvThreads = 256;
constexpr int kTsvSeg = 32;                                   // bytes per thread
constexpr int kTsvChunk = kTsvThreads * kTsvSeg;              // bytes per block
constexpr int kMaxTsvFiles = 64;
constexpr unsigned kTsvUnhandled = 1u, kTsvOverflow = 2u, kTsvNan0 = 4u, kTsvNan1 = 8u;     // status bits per file
constexpr unsigned kTsvBad = kTsvUnhandled | kTsvOverflow;

struct TsvFile {
    const char *text;          // 16-byte aligned base of the file's bytes on the device
    int64_t begin, end;        // data region [begin, end): the host has skipped header rows
    double2 *out;              // rows land here
    int64_t cap_rows;
};
struct TsvWave { TsvFile f[kMaxTsvFiles]; };

// The text as a kernel sees it: a window [lo, hi) of it staged in shared memory (k_tsv_parse) with global memory behind
// it, or global memory only (k_tsv_count: every byte is read once, staging would not pay).
struct TsvText {
    const char *g;             // global: g[i] for any i in the file
    const char *s;             // s[i] valid for lo <= i < hi (shared window, biased by -lo); null: no window
    int64_t lo, hi;
    __device__ __forceinline__ char at(int64_t i) const { return (s && i >= lo && i < hi) ? s[i] : g[i]; }
    __device__ __forceinline__ const char *span(int64_t b, int64_t e) const {      // contiguous [b, e)
        return (s && b >= lo && e <= hi) ? s + b : g + b;
    }
    __device__ __forceinline__ uint4 load16(int64_t i) const {                     // i multiple of 16
        return (s && i >= lo && i + 16 <= hi) ? *reinterpret_cast<const uint4 *>(s + i)
                                              : *reinterpret_cast<const uint4 *>(g + i);
    }
};

// a data line: first non-blank byte is neither the end of the line nor '#'
__device__ __forceinline__ bool tsv_is_data_line(const TsvText &t, int64_t i, int64_t end) {
    char c = 0;
    while (i < end && ((c = t.at(i)) == ' ' || c == '\r' || c == '\t')) ++i;
    return i < end && c != '\n' && c != '#';
}

// bit k of the result: byte (seg0 + k) starts a data line
__device__ __forceinline__ uint32_t tsv_line_starts(const TsvFile &f, const TsvText &t, int64_t seg0) {
    if (seg0 >= f.end || seg0 + kTsvSeg <= f.begin) return 0u;
    const uint4 a = t.load16(seg0);
    const uint4 b = t.load16(seg0 + 16);                      // the text buffer is padded to 32 bytes
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t nl = 0u;                                         // bit k: byte k is '\n'
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t x = w[k] ^ 0x0a0a0a0au;                // zero bytes where '\n'
        const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;
        nl |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    uint32_t starts = nl << 1;                                // a line starts after every '\n' ...
    const bool prev_nl = (seg0 == f.begin) || (seg0 > f.begin && t.at(seg0 - 1) == '\n');
    if (prev_nl) starts |= 1u;                                // ... and at the first byte of the data region
    // only positions inside [begin, end)
    if (seg0 < f.begin) starts &= ~0u << (int)(f.begin - seg0), starts |= 1u << (int)(f.begin - seg0);
    if (f.end - seg0 < kTsvSeg) starts &= (1u << (int)(f.end - seg0)) - 1u;
    uint32_t data = 0u;
    for (uint32_t m = starts; m; m &= m - 1u) {
        const int k = __ffs(m) - 1;
        if (tsv_is_data_line(t, seg0 + k, f.end)) data |= 1u << k;
    }
    return data;
}

__global__ void __launch_bounds__(kTsvThreads)
k_tsv_count(const __grid_constant__ TsvWave wv, int chunks_max, uint32_t *__restrict__ chunkrows) {
    const TsvFile &f = wv.f[blockIdx.y];
    const int64_t chunk0 = (f.begin / kTsvChunk + blockIdx.x) * (int64_t)kTsvChunk;     // chunks are aligned to 8 KB
    __shared__ uint32_t s_w[kTsvThreads / 32];
    uint32_t n = 0u;
    const TsvText t{f.text, nullptr, 0, 0};
    if (chunk0 < f.end) n = __popc(tsv_line_starts(f, t, chunk0 + (int64_t)threadIdx.x * kTsvSeg));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0u;
#pragma unroll
        for (int k = 0; k < kTsvThreads / 32; ++k) t += s_w[k];
        chunkrows[(size_t)blockIdx.y * chunks_max + blockIdx.x] = t;
    }
}

// grid (files), 1024 threads: chunkrows[f][0..chunks_max) -> exclusive prefix; rows[f] = total
__global__ void __launch_bounds__(1024)
k_tsv_scan(int chunks_max, uint32_t *__restrict__ chunkrows, int64_t *__restrict__ rows) {
    uint32_t *c = chunkrows + (size_t)blockIdx.x * chunks_max;
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_run;
    if (threadIdx.x == 0) s_run = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < chunks_max; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < chunks_max ? c[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        if (w == 0) {
            const uint32_t x = s_w[lane];
            uint32_t ix = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ix, o);
                if (lane >= o) ix += t;
            }
            s_w[lane] = ix - x;
        }
        __syncthreads();
        const uint32_t run = s_run;
        if (i < chunks_max) c[i] = run + s_w[w] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + s_w[w] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) rows[blockIdx.x] = (int64_t)s_run;
}

// One pass over a line, one byte per step, every lane of a warp in step: each tab-separated field is read as a decimal
// while it goes by ([+-]digits[.digits][e[+-]digits] between blanks; the same grammar as kdejs_decimal::parse_field and
// the host reader's trimming of ' ' and '\r'), and only the last two fields' results are kept -- id columns simply
// come out "not a number" and are shifted away.  The first version found the line's end, walked back to the last two
// tabs, trimmed and then parsed each field: four loops of different length per lane (ncu: 12 of 32 lanes active, 230
// warp instructions per line).
struct TsvField {
    unsigned long long w;      // significant digits
    int q;                     // decimal exponent: value = w * 10^q
    int kind;                  // 0 number, 1 empty (missing value: NaN), 2 not covered by the device parser
    bool neg;
};
enum { kFsStart, kFsSign, kFsInt, kFsFrac, kFsExpMark, kFsExpSign, kFsExp, kFsTrail, kFsBad };

struct TsvFieldState {
    unsigned long long w;
    int state, sig, frac, ndig, ex, exdig;
    bool neg, eneg;
    __device__ __forceinline__ void reset() { w = 0ull; state = kFsStart; sig = frac = ndig = ex = exdig = 0; neg = eneg = false; }
    __device__ __forceinline__ void digit(unsigned d, bool fraction) {
        ++ndig;
        if (fraction) ++frac;
        if (sig == 0 && d == 0u) return;                      // leading zeros
        if (++sig > 19) { state = kFsBad; return; }
        w = w * 10ull + d;
    }
    __device__ __forceinline__ void step(char c) {
        const unsigned d = (unsigned)(c - '0');
        if (state == kFsBad) return;
        if (d < 10u) {
            if (state <= kFsInt) { state = kFsInt; digit(d, false); }
            else if (state == kFsFrac) digit(d, true);
            else if (state <= kFsExp) { state = kFsExp; if (++exdig > 4) state = kFsBad; else ex = ex * 10 + (int)d; }
            else state = kFsBad;                              // a digit after trailing blanks
        } else if (c == ' ' || c == '\r') {
            if (state == kFsStart || state == kFsTrail) return;
            state = (state == kFsInt || state == kFsFrac || state == kFsExp) ? kFsTrail : kFsBad;
        } else if (c == '+' || c == '-') {
            if (state == kFsStart) { state = kFsSign; neg = (c == '-'); }
            else if (state == kFsExpMark) { state = kFsExpSign; eneg = (c == '-'); }
            else state = kFsBad;
        } else if (c == '.') {
            state = (state <= kFsInt) ? kFsFrac : kFsBad;
        } else if (c == 'e' || c == 'E') {
            state = ((state == kFsInt || state == kFsFrac) && ndig > 0) ? kFsExpMark : kFsBad;
        } else state = kFsBad;
    }
    __device__ __forceinline__ TsvField finish() const {
        TsvField r{w, 0, 2, neg};
        if (state == kFsStart) { r.kind = 1; return r; }      // nothing but blanks: a missing value
        // where the field may end: after digits (integer / fraction / exponent), possibly followed by blanks
        const bool ended_ok = state == kFsInt || state == kFsFrac || state == kFsExp || state == kFsTrail;
        if (!ended_ok || ndig == 0) return r;
        const int qq = (eneg ? -ex : ex) - frac;
        if (w != 0ull && (qq < -27 || qq > 27)) return r;
        r.q = qq;
        r.kind = 0;
        return r;
    }
};

__device__ __forceinline__ bool tsv_field_value(const TsvField &fl, double &v) {
    if (fl.kind == 1) { v = __longlong_as_double(0x7ff8000000000000ll); return true; }
    if (fl.kind != 0) return false;
    const double r = fl.w ? kdejs_decimal::scale_exact(fl.w, fl.q) : 0.0;
    v = fl.neg ? -r : r;
    return true;
}

// order-preserving map double -> u64 (NaN excluded by the caller): bounds are folded with integer atomicMax
__device__ __forceinline__ unsigned long long tsv_key(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

constexpr int kTsvTail = 1024;                                // bytes staged beyond a chunk: a line that starts in the
                                                              // chunk and ends within them is parsed from shared memory
// bkeys[file][4]: atomicMax of ~key(min col 0), key(max col 0), ~key(min col 1), key(max col 1); zero before the call
__global__ void __launch_bounds__(kTsvThreads)
k_tsv_parse(const __grid_constant__ TsvWave wv, int chunks_max, const uint32_t *__restrict__ chunkrows,
            unsigned *__restrict__ status, unsigned long long *__restrict__ bkeys) {
    const TsvFile &f = wv.f[blockIdx.y];
    const int64_t chunk0 = (f.begin / kTsvChunk + blockIdx.x) * (int64_t)kTsvChunk;
    if (chunk0 >= f.end) return;
    // the chunk, 16 bytes before it (the byte that decides whether its first byte starts a line) and kTsvTail after
    __shared__ __align__(16) char s_text[16 + kTsvChunk + kTsvTail];
    const int64_t w0 = chunk0 >= 16 ? chunk0 - 16 : 0;
    const int64_t w1 = min(w0 + (int64_t)sizeof s_text, (f.end + 15) / 16 * 16);         // the buffer is padded past `end`
    for (int64_t i = w0 + (int64_t)threadIdx.x * 16; i < w1; i += kTsvThreads * 16)
        *reinterpret_cast<uint4 *>(s_text + (i - w0)) = *reinterpret_cast<const uint4 *>(f.text + i);
    __syncthreads();
    const TsvText t{f.text, s_text - w0, w0, w1};
    const int64_t seg0 = chunk0 + (int64_t)threadIdx.x * kTsvSeg;
    const uint32_t starts = tsv_line_starts(f, t, seg0);
    const uint32_t n = __popc(starts);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    __shared__ uint32_t s_w[kTsvThreads / 32];
    __shared__ unsigned long long s_k[4][kTsvThreads / 32];
    // The chunk's line starts, compacted in file order: a thread that finds two starts in its 32 bytes and one that
    // finds none would otherwise leave most lanes of a warp idle (ncu: 8.75 of 32 lanes active per instruction).
    // A data line holds at least one byte and a newline: at most kTsvChunk / 2 of them start in a chunk.
    __shared__ uint16_t s_line[kTsvChunk / 2];
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    uint32_t before = 0u, total = 0u;
#pragma unroll
    for (int k = 0; k < kTsvThreads / 32; ++k) {
        if (k < w) before += s_w[k];
        total += s_w[k];
    }
    {
        uint32_t slot = before + inc - n;
        for (uint32_t m = starts; m; m &= m - 1u) s_line[slot++] = (uint16_t)(threadIdx.x * kTsvSeg + __ffs(m) - 1);
    }
    __syncthreads();
    const int64_t row0 = (int64_t)chunkrows[(size_t)blockIdx.y * chunks_max + blockIdx.x];
    unsigned flags = 0u;
    unsigned long long k4[4] = {0ull, 0ull, 0ull, 0ull};
    const uint32_t wsize = (uint32_t)(w1 - w0);
    for (uint32_t i = threadIdx.x; i < total; i += kTsvThreads) {
        const int64_t b = chunk0 + s_line[i];
        uint32_t o = (uint32_t)(b - w0);                      // offset in the staged window
        const uint32_t o_end = (uint32_t)min(f.end - w0, (int64_t)0x7fffffff);
        TsvFieldState cur;
        cur.reset();
        TsvField prev{0ull, 0, 2, false};
        int tabs = 0;
        for (; o < o_end; ++o) {
            const char c = o < wsize ? s_text[o] : f.text[w0 + o];      // (a line longer than the window runs on in global memory)
            if (c == '\n') break;
            if (c == '\t') { prev = cur.finish(); cur.reset(); ++tabs; }
            else cur.step(c);
        }
        double x = 0.0, y = 0.0;
        const TsvField last = cur.finish();
        if (tabs == 0 || !tsv_field_value(prev, x) || !tsv_field_value(last, y)) { flags |= kTsvUnhandled; x = y = 0.0; }
        const int64_t row = row0 + i;
        if (row < f.cap_rows) f.out[row] = make_double2(x, y);
        else flags |= kTsvOverflow;
        if (x != x) flags |= kTsvNan0; else { const unsigned long long k = tsv_key(x); k4[0] = max(k4[0], ~k); k4[1] = max(k4[1], k); }
        if (y != y) flags |= kTsvNan1; else { const unsigned long long k = tsv_key(y); k4[2] = max(k4[2], ~k); k4[3] = max(k4[3], k); }
    }
    if (bkeys) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) k4[c] = max(k4[c], __shfl_xor_sync(0xffffffffu, k4[c], o));
            if (lane == 0) s_k[c][w] = k4[c];
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            unsigned long long k = 0ull;
#pragma unroll
            for (int j = 0; j < kTsvThreads / 32; ++j) k = max(k, s_k[threadIdx.x][j]);
            if (k) atomicMax(&bkeys[(size_t)blockIdx.y * 4 + threadIdx.x], k);
        }
    }
    if (flags) atomicOr(&status[blockIdx.y], flags);
}

}  // namespace kdejs
