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
//   k_tsv_parse   every line start parses its line and writes row (x, y) to its place
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
constexpr unsigned kTsvUnhandled = 1u, kTsvOverflow = 2u;     // status bits per file

struct TsvFile {
    const char *text;          // 16-byte aligned base of the file's bytes on the device
    int64_t begin, end;        // data region [begin, end): the host has skipped header rows
    double2 *out;              // rows land here
    int64_t cap_rows;
};
struct TsvWave { TsvFile f[kMaxTsvFiles]; };

// a data line: first non-blank byte is neither the end of the line nor '#'
__device__ __forceinline__ bool tsv_is_data_line(const char *t, int64_t i, int64_t end) {
    while (i < end && (t[i] == ' ' || t[i] == '\r' || t[i] == '\t')) ++i;
    return i < end && t[i] != '\n' && t[i] != '#';
}

// bit k of the result: byte (seg0 + k) starts a data line
__device__ __forceinline__ uint32_t tsv_line_starts(const TsvFile &f, int64_t seg0) {
    if (seg0 >= f.end || seg0 + kTsvSeg <= f.begin) return 0u;
    const uint4 a = *reinterpret_cast<const uint4 *>(f.text + seg0);
    const uint4 b = *reinterpret_cast<const uint4 *>(f.text + seg0 + 16);      // the text buffer is padded to 32 bytes
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t nl = 0u;                                         // bit k: byte k is '\n'
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t x = w[k] ^ 0x0a0a0a0au;                // zero bytes where '\n'
        const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;
        nl |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    uint32_t starts = nl << 1;                                // a line starts after every '\n' ...
    const bool prev_nl = (seg0 == f.begin) || (seg0 > f.begin && f.text[seg0 - 1] == '\n');
    if (prev_nl) starts |= 1u;                                // ... and at the first byte of the data region
    // only positions inside [begin, end)
    if (seg0 < f.begin) starts &= ~0u << (int)(f.begin - seg0), starts |= 1u << (int)(f.begin - seg0);
    if (f.end - seg0 < kTsvSeg) starts &= (1u << (int)(f.end - seg0)) - 1u;
    uint32_t data = 0u;
    for (uint32_t m = starts; m; m &= m - 1u) {
        const int k = __ffs(m) - 1;
        if (tsv_is_data_line(f.text, seg0 + k, f.end)) data |= 1u << k;
    }
    return data;
}

__global__ void __launch_bounds__(kTsvThreads)
k_tsv_count(const __grid_constant__ TsvWave wv, int chunks_max, uint32_t *__restrict__ chunkrows) {
    const TsvFile &f = wv.f[blockIdx.y];
    const int64_t chunk0 = (f.begin / kTsvChunk + blockIdx.x) * (int64_t)kTsvChunk;     // chunks are aligned to 8 KB
    __shared__ uint32_t s_w[kTsvThreads / 32];
    uint32_t n = 0u;
    if (chunk0 < f.end) n = __popc(tsv_line_starts(f, chunk0 + (int64_t)threadIdx.x * kTsvSeg));
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

// one line [b, e): its last two tab-separated fields
__device__ __forceinline__ bool tsv_parse_line(const char *t, int64_t b, int64_t e, double &x, double &y) {
    int64_t p = e, t2 = -1, s1 = b;
    while (p > b) {
        --p;
        if (t[p] == '\t') {
            if (t2 < 0) t2 = p;
            else { s1 = p + 1; break; }
        }
    }
    if (t2 < 0) return false;
    int64_t fb[2] = {s1, t2 + 1}, fe[2] = {t2, e};
    double v[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        while (fb[k] < fe[k] && (t[fb[k]] == ' ' || t[fb[k]] == '\r')) ++fb[k];
        while (fe[k] > fb[k] && (t[fe[k] - 1] == ' ' || t[fe[k] - 1] == '\r')) --fe[k];
        if (fb[k] >= fe[k]) v[k] = __longlong_as_double(0x7ff8000000000000ll);          // empty field: missing
        else if (!kdejs_decimal::parse_field(t + fb[k], t + fe[k], v[k])) return false;
    }
    x = v[0];
    y = v[1];
    return true;
}

__global__ void __launch_bounds__(kTsvThreads)
k_tsv_parse(const __grid_constant__ TsvWave wv, int chunks_max, const uint32_t *__restrict__ chunkrows,
            unsigned *__restrict__ status) {
    const TsvFile &f = wv.f[blockIdx.y];
    const int64_t chunk0 = (f.begin / kTsvChunk + blockIdx.x) * (int64_t)kTsvChunk;
    if (chunk0 >= f.end) return;
    const int64_t seg0 = chunk0 + (int64_t)threadIdx.x * kTsvSeg;
    const uint32_t starts = tsv_line_starts(f, seg0);
    const uint32_t n = __popc(starts);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __shared__ uint32_t s_w[kTsvThreads / 32];
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    uint32_t before = 0u;
    for (int k = 0; k < w; ++k) before += s_w[k];
    int64_t row = (int64_t)chunkrows[(size_t)blockIdx.y * chunks_max + blockIdx.x] + before + inc - n;
    unsigned flags = 0u;
    for (uint32_t m = starts; m; m &= m - 1u, ++row) {
        const int64_t b = seg0 + __ffs(m) - 1;
        int64_t e = b;
        while (e < f.end && f.text[e] != '\n') ++e;
        double x, y;
        if (!tsv_parse_line(f.text, b, e, x, y)) { flags |= kTsvUnhandled; x = y = 0.0; }
        if (row < f.cap_rows) f.out[row] = make_double2(x, y);
        else flags |= kTsvOverflow;
    }
    if (flags) atomicOr(&status[blockIdx.y], flags);
}

// grid (files), 256 threads: (min, max) of both columns of a parsed table, as np.min / np.max give them (NaN when the
// column holds one) -- the bounds scripts/distance_to_gridsearch.py:112-125 tests before it keeps a table as best.
__global__ void __launch_bounds__(256)
k_tsv_bounds(const __grid_constant__ TsvWave wv, double *__restrict__ bounds /*[file][4]*/) {
    const TsvFile &f = wv.f[blockIdx.x];
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double lo0 = inf, hi0 = -inf, lo1 = inf, hi1 = -inf;
    bool nan0 = false, nan1 = false;
    for (int64_t i = threadIdx.x; i < f.cap_rows; i += blockDim.x) {
        const double2 v = f.out[i];
        nan0 |= (v.x != v.x);
        nan1 |= (v.y != v.y);
        lo0 = fmin(lo0, v.x); hi0 = fmax(hi0, v.x);         // fmin / fmax ignore NaN
        lo1 = fmin(lo1, v.y); hi1 = fmax(hi1, v.y);
    }
    __shared__ double s_v[4][256];
    __shared__ int s_nan[2];
    if (threadIdx.x < 2) s_nan[threadIdx.x] = 0;
    __syncthreads();
    if (nan0) s_nan[0] = 1;
    if (nan1) s_nan[1] = 1;
    s_v[0][threadIdx.x] = lo0; s_v[1][threadIdx.x] = hi0; s_v[2][threadIdx.x] = lo1; s_v[3][threadIdx.x] = hi1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_v[0][threadIdx.x] = fmin(s_v[0][threadIdx.x], s_v[0][threadIdx.x + o]);
            s_v[1][threadIdx.x] = fmax(s_v[1][threadIdx.x], s_v[1][threadIdx.x + o]);
            s_v[2][threadIdx.x] = fmin(s_v[2][threadIdx.x], s_v[2][threadIdx.x + o]);
            s_v[3][threadIdx.x] = fmax(s_v[3][threadIdx.x], s_v[3][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) {
        const double nanv = __longlong_as_double(0x7ff8000000000000ll);
        bounds[(size_t)blockIdx.x * 4 + threadIdx.x] = s_nan[threadIdx.x >> 1] ? nanv : s_v[threadIdx.x][0];
    }
}

}  // namespace kdejs
