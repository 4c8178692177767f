// kdejs_kernels.cuh -- hand-written sm_100a kernels for PopPUNK-mod's KDE discrepancy.
//
// One evaluation = KDE_JS_divergence(df1, df2, gamma, eps, log)  (/root/reference/KDE_distance.py:104-130).
// A "wave" is up to kMaxWaveItems independent evaluations processed by the same eight launches:
//
//   k_minmax         joint NaN-ignoring min/max per raw point set (MinMaxScaler.fit, KDE_distance.py:81-82)
//   k_scale_count    X*scale_+min_, NaN->0 (KDE_distance.py:85-86, :53): cell id and in-warp rank per point,
//                    per-warp / per-block cell histograms
//   k_scan_*         exclusive scan of the (cell, block) counts -> cell starts and block bases
//   k_cell_moments   (fine grids, dense cells) five moments per cell for cells wholly inside a tile's support
//   k_place          stable counting-sort placement of the scaled points by cell (pure scatter)
//   k_tile_classify  candidates per 4x4-node tile -> work list in ten weight classes; empty tiles finished here
//   k_density_*      Epanechnikov sums on the R x R grid (get_kde + generate_samples, :48-64),
//                    **gamma + eps and per-tile normaliser partials (:71-76)
//   k_set_totals     per-set normaliser (z.sum(), :76)
//   k_js             normalise, Jensen-Shannon distance (scipy jensenshannon, :129)
//
// Everything is fp64 unless a kernel says otherwise; every reduction has a fixed shape, so results
// are bit-reproducible run to run and independent of batch composition.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace kdejs {

constexpr int kMaxWaveItems = 64;
constexpr int kMaxWaveSets = 2 * kMaxWaveItems;
constexpr int kMaxCellsSide = 128;                // cells per side of the culling grid (64 unless the grid is fine)
constexpr int kMaxCells = kMaxCellsSide * kMaxCellsSide;
constexpr int kSortWarps = 4;
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kTile = 4;                          // nodes per tile side (16 accumulators per thread)
constexpr int kDensThreads = 128;
constexpr int kMaxRows = 32;                      // cell rows a tile may have to visit
constexpr int kNumClasses = 10;
// The sets of a wave are split into up to kMaxGroups groups whose sorted points fit the L2 together; the list is
// ordered (group, class), so the ~36 tiles that share a cell's points run close in time and the points are read
// from DRAM about once, while heaviest-first order still holds inside a group and a group's light tail overlaps
// the next group's heavy tiles (one persistent launch, one ticket).
constexpr int kMaxGroups = 4;      // measured: 4 groups bring the kernel's DRAM reads from 606 MB to 317 MB per 64
                                   // evaluations (algorithmic 339 MB incl. writes); 8 cost 1 % in the list lookup
constexpr int kNumLists = kMaxGroups * kNumClasses;
constexpr int kTicketA = kNumLists, kTicketB = kNumLists + 1;
constexpr int kSplitParts = kNumLists + 2;        // part slots handed out to split tiles (see SplitBufs)
constexpr int kNumCounters = kNumLists + 3;
constexpr int kRedoBase = kNumCounters;           // counters[kRedoBase ...]: the same layout again for the redo list
constexpr int kAllCounters = 2 * kNumCounters;
constexpr int kJsThreads = 256;
constexpr int kJsNodesPerBlock = 2048;
constexpr int kMinmaxThreads = 256;

// Very heavy tiles are split.  A block-per-tile item (class 0) with >= 2^kSplitLog2 candidates is entered K times
// into its list, K = ceil(candidates / 2^kPartLog2) <= kMaxParts; part k walks chunks [gtot k / K, gtot (k+1) / K)
// of the tile's flattened sequence of 32-candidate chunks and leaves its 16 partial node sums in `partial`; the
// part that finishes last (ticket) adds the K partials in part order and runs the tile's epilogue.  The split is
// a function of the tile's candidate runs alone, so results stay bit-reproducible and identical between one GPU
// and a band of a sharded evaluation.  Why: on config C5 the heaviest tiles hold 1.7 M candidates -- 0.9 ms on
// one block, more than a rank's whole share of the density work once the grid is spread over eight GPUs.
constexpr int kSplitLog2 = 17, kPartLog2 = 16, kMaxParts = 32;
struct SplitBufs {
    uint2 *partinfo;       // per class-0 list entry [group][capacity]: {part | parts << 8, first part slot}
    double *partial;       // [part slot][16]
    unsigned *ticket;      // [part slot] (the tile's first slot is used), zero between waves (self-resetting)
    unsigned parts_cap;    // part slots available; a tile that would not fit is not split
};

struct RawSet {            // a point set as the caller gave it: (n,2) interleaved fp64, device memory
    const double *p;
    int64_t n;
};
struct SetDesc {           // one KDE to build: `rs_self` scaled jointly with `rs_other`
    int32_t rs_self, rs_other, n, nb;   // nb = blocks that own a chunk of it in the sort
    int64_t off;                        // first point in the wave's scaled/sorted arrays
    double knorm_over_n;                // kernel normaliser / n: density = sum * knorm_over_n
};
struct WaveDesc {          // passed by value as a __grid_constant__ kernel parameter (6 KB; CUDA >= 12.1 allows 32 KB)
    RawSet raw[kMaxWaveSets];
    SetDesc set[kMaxWaveSets];
    int32_t n_raw, n_sets;
};
struct RawStat {           // NaN-ignoring column extrema of one raw set (after the optional log)
    double mn[2], mx[2];
    int32_t flags, pad;    // bit 0: a non-NaN infinity is present
};
struct GridGeom {
    int32_t R, NT, C, gamma_mode;        // NT tiles per side, C cells per side; gamma_mode 1: gamma==1
    double gmin, step, h, inv_h, dh;     // dh = step/h
    double margin, cell_x0, cell_invw;   // cell(x) = clamp(floor((x-cell_x0)*cell_invw), 0, C-1)
    double cell_w;                       // nominal cell side = extent / C
    double knorm, gamma, eps;
    int32_t tile_row_begin, tile_row_end;   // tile rows (index of column 0) this launch covers
    int32_t tile_col_begin, tile_col_end;   // ... and tile columns (index of column 1): band mode shards along these,
                                            // because cell ROWS (the sort's major order) run along column 1
    int32_t sets_per_group, split_cfg;      // work-list locality: sets [k*spg, (k+1)*spg) form group k (see kMaxGroups);
                                            // split_cfg = split_log2 | part_log2 << 8 (very heavy tiles, see SplitBufs)
    int32_t heavy_log2, use_moments;        // tiles with >= 2^heavy_log2 candidates are processed per block;
                                            // use_moments: cells wholly inside a tile's support go in as moments
    // fp32 path (k_density_binned32): sorted points are also kept as 32-bit fixed point in units of h * 2^-q_shift
    // (modulo 2^32: only differences to a tile's first node are ever used, and those stay below 2^31)
    int32_t q_shift, q_pad;                 // 0: geometry not eligible (reach too long for the format), fp64 path is used
    double q_scale, q_lo, q_hi;             // q = (u32)(i64)rint((clamp(x, q_lo, q_hi) - gmin) * q_scale)
    float q_unit, q_dh;                     // 2^-q_shift; dh as float
};

__device__ __forceinline__ double pre_transform(double v, int log_flag, double eps) {
    return log_flag ? log(v + eps) : v;    // KDE_distance.py:105-106
}
__device__ __forceinline__ int cell_of(double x, double x0, double invw, int C) {
    // floor, then clamp: monotone in x (the tile side below relies on it).  __double2int_rd saturates for
    // out-of-range values and maps NaN to 0, so the integer clamp covers every input.
    return min(max(__double2int_rd((x - x0) * invw), 0), C - 1);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Fixed-shape block sum (blockDim.x multiple of 32, <= 1024); result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double *s_warp /*[32]*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_warp[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_warp[i];
    return t;
}

// ------------------------------------------------------------------------------------------
// k_minmax: grid (nblk, n_raw).  Last block per raw set (ticket) folds the block partials.
// min/max are exact in any order, so the result does not depend on the launch shape.
__global__ void __launch_bounds__(kMinmaxThreads)
k_minmax(const __grid_constant__ WaveDesc wd, int log_flag, double eps,
         RawStat *__restrict__ part, unsigned *__restrict__ ticket, RawStat *__restrict__ out) {
    const int r = blockIdx.y;
    const double2 *__restrict__ p = reinterpret_cast<const double2 *>(wd.raw[r].p);
    const int64_t n = wd.raw[r].n;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double mn0 = INF, mx0 = -INF, mn1 = INF, mx1 = -INF;
    int flags = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                  // four independent loads in flight per thread
            const int64_t i = i0 + u * stride;
            v[u] = (i < n) ? p[i] : make_double2(__longlong_as_double(0x7ff8000000000000LL),
                                                 __longlong_as_double(0x7ff8000000000000LL));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double x = pre_transform(v[u].x, log_flag, eps), y = pre_transform(v[u].y, log_flag, eps);
            if (x == x) { mn0 = fmin(mn0, x); mx0 = fmax(mx0, x); flags |= isinf(x) ? 1 : 0; }
            if (y == y) { mn1 = fmin(mn1, y); mx1 = fmax(mx1, y); flags |= isinf(y) ? 1 : 0; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn0 = fmin(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
        mx0 = fmax(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
        mn1 = fmin(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
        mx1 = fmax(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        flags |= __shfl_xor_sync(0xffffffffu, flags, o);
    }
    __shared__ RawStat s_part[kMinmaxThreads / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_part[w].mn[0] = mn0; s_part[w].mx[0] = mx0; s_part[w].mn[1] = mn1;
                     s_part[w].mx[1] = mx1; s_part[w].flags = flags; }
    __syncthreads();
    if (threadIdx.x == 0) {
        RawStat a = s_part[0];
        for (int i = 1; i < kMinmaxThreads / 32; ++i) {
            a.mn[0] = fmin(a.mn[0], s_part[i].mn[0]); a.mx[0] = fmax(a.mx[0], s_part[i].mx[0]);
            a.mn[1] = fmin(a.mn[1], s_part[i].mn[1]); a.mx[1] = fmax(a.mx[1], s_part[i].mx[1]);
            a.flags |= s_part[i].flags;
        }
        a.pad = 0;
        part[(size_t)r * gridDim.x + blockIdx.x] = a;
        __threadfence();
        unsigned t = atomicAdd(&ticket[r], 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {          // warp 0 of the last block folds the block partials
        __threadfence();
        const volatile RawStat *vp = part + (size_t)r * gridDim.x;
        double a0 = INF, b0 = -INF, a1 = INF, b1 = -INF;
        int fl = 0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) {
            a0 = fmin(a0, vp[i].mn[0]); b0 = fmax(b0, vp[i].mx[0]);
            a1 = fmin(a1, vp[i].mn[1]); b1 = fmax(b1, vp[i].mx[1]);
            fl |= vp[i].flags;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 = fmin(a0, __shfl_xor_sync(0xffffffffu, a0, o));
            b0 = fmax(b0, __shfl_xor_sync(0xffffffffu, b0, o));
            a1 = fmin(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            b1 = fmax(b1, __shfl_xor_sync(0xffffffffu, b1, o));
            fl |= __shfl_xor_sync(0xffffffffu, fl, o);
        }
        if (threadIdx.x == 0) {
            RawStat a;
            a.mn[0] = a0; a.mx[0] = b0; a.mn[1] = a1; a.mx[1] = b1; a.flags = fl; a.pad = 0;
            out[r] = a;
            ticket[r] = 0;     // self-resetting: the next wave needs no memset
        }
    }
}

// sklearn MinMaxScaler parameters from the joint extrema of two raw sets
// (sklearn:preprocessing/_data.py:476-545; _handle_zeros_in_scale :101-133).
__device__ __forceinline__ void joint_scaler(const RawStat &a, const RawStat &b, int col,
                                             double &scale, double &min_) {
    double mn = fmin(a.mn[col], b.mn[col]), mx = fmax(a.mx[col], b.mx[col]);
    double range = mx - mn;
    if (range < 10.0 * 2.220446049250313e-16) range = 1.0;
    scale = 1.0 / range;
    min_ = 0.0 - __dmul_rn(mn, scale);
}

// ------------------------------------------------------------------------------------------
// The scaler of one set (joint with its partner set) and the transform of one raw point:
// X *= scale_; X += min_  as two separately rounded operations (never an FMA), then NaN -> 0.
struct Scaler { double sx, mx, sy, my; };
__device__ __forceinline__ Scaler make_scaler(const RawStat *__restrict__ rstat, const SetDesc &sd, int identity) {
    Scaler sc;
    if (identity) { sc.sx = 1.0; sc.mx = 0.0; sc.sy = 1.0; sc.my = 0.0; }
    else {
        joint_scaler(rstat[sd.rs_self], rstat[sd.rs_other], 0, sc.sx, sc.mx);
        joint_scaler(rstat[sd.rs_self], rstat[sd.rs_other], 1, sc.sy, sc.my);
    }
    return sc;
}
__device__ __forceinline__ double2 scale_point(double2 v, const Scaler &sc, int log_flag, double eps, int identity) {
    double X = __dadd_rn(__dmul_rn(pre_transform(v.x, log_flag, eps), sc.sx), sc.mx);
    double Y = __dadd_rn(__dmul_rn(pre_transform(v.y, log_flag, eps), sc.sy), sc.my);
    if (!identity) {                  // KDE_distance.py:53
        if (X != X) X = 0.0;
        if (Y != Y) Y = 0.0;
    }
    return make_double2(X, Y);
}

// Zero n u16 counters in dynamic shared memory (16-byte aligned) with 128-bit stores.
__device__ __forceinline__ void zero_u16(uint16_t *p, int n) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    const int n16 = n / 8;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) q[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = n16 * 8 + threadIdx.x; i < n; i += blockDim.x) p[i] = 0;
}

constexpr int kSortUnroll = 4;        // independent 32-point groups a warp keeps in flight
// How the lanes of a 32-point group find the same-cell lanes below them (k_scale_count):
//   0  MATCH.ANY on the cell id.  Its cost grows with the number of DISTINCT values in the warp (tools/ubench/
//      match_probe.cu on a B200 at this kernel's 24 warps per SM: 2 SM-cycles per warp-level MATCH with one value,
//      21 with 16, 47 with 32 distinct ones) and unsorted points land in ~30 different cells per group: half of the
//      kernel's 94 SM-cycles per group.
//   1  one ballot per bit of the cell id (26 SM-cycles in the probe, but slower in the kernel: 2.06 -> 2.92 us
//      per evaluation, measured in rounds 1 and 2).
//   2  a 15-step shuffle network sorts (cell, lane) keys across the warp; same-cell lanes become neighbours in lane
//      order and a ballot of the segment heads gives every point its rank (17 SM-cycles in the probe) -- and yet
//      slower in the kernel as well: 2.12 -> 2.50 us per evaluation, 0.078 -> 0.093 ms for a band of config C5
//      (same results bit for bit).  What the kernel waits for is the MIO queue as a whole -- MATCH, the histogram's
//      LDS/STS and the shuffles share it -- so trading one MATCH for sixteen shuffles does not pay.
// 0 is the default; 1 and 2 stay as compile-time experiments (-DKDEJS_RANK_MODE=...).
#ifndef KDEJS_RANK_MODE
#define KDEJS_RANK_MODE 0
#endif
constexpr int kRankMode = KDEJS_RANK_MODE;
constexpr int kMaxChw = 8192;         // points per warp sub-chunk: 4 warps * kMaxChw < 65536 (u16 counters)

// k_scale_count: grid (nb_max, n_sets), kSortThreads threads, dyn smem kSortWarps*ncells u16 (a warp's
// sub-chunk holds at most kMaxChw points, so 16-bit counters cannot overflow).
// Block b owns points [b*chw*kSortWarps, ...), warp w the w-th sub-chunk: index order == (block,
// warp, lane) order.  For every point it writes the cell id and the point's RANK among the same-cell
// points of its warp sub-chunk (earlier iterations + lower lanes); per block it writes the cell
// histogram and, per warp, the number of same-cell points in earlier warps of the block.  With the
// block bases from k_scan these three numbers are the point's stable sorted position, so k_place needs
// no ranking of its own.  The scaled coordinates are written only when `scaled` is non-null (dense
// path) -- the binned path recomputes them in k_place.
__global__ void __launch_bounds__(kSortThreads)
k_scale_count(const __grid_constant__ WaveDesc wd, const RawStat *__restrict__ rstat,
              int log_flag, double eps, int identity_scaler, int C, double cell_x0,
              double cell_invw, int chw, int nb_max, double2 *__restrict__ scaled,
              uint32_t *__restrict__ cellrank,
              uint16_t *__restrict__ warpprefix, uint32_t *__restrict__ blockhist) {
    extern __shared__ uint16_t s_hist[];
    const SetDesc sd = wd.set[blockIdx.y];
    if ((int)blockIdx.x >= sd.nb) return;
    const int ncells = C * C;
    const int cell_bits = 32 - __clz(max(ncells - 1, 1));
    zero_u16(s_hist, kSortWarps * ncells);
    __syncthreads();
    const Scaler sc = make_scaler(rstat, sd, identity_scaler);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double2 *__restrict__ raw = reinterpret_cast<const double2 *>(wd.raw[sd.rs_self].p);
    const int64_t beg = ((int64_t)blockIdx.x * kSortWarps + warp) * chw;
    const int64_t end = min(beg + (int64_t)chw, (int64_t)sd.n);
    uint16_t *hw = s_hist + warp * ncells;
    for (int64_t j0 = beg; j0 < end; j0 += 32 * kSortUnroll) {
        double2 v[kSortUnroll];
#pragma unroll
        for (int u = 0; u < kSortUnroll; ++u) {
            const int64_t j = j0 + u * 32 + lane;
            v[u] = (j < end) ? raw[j] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < kSortUnroll; ++u) {
            const int64_t j = j0 + u * 32 + lane;
            const bool valid = j < end;
            int cell = 0;
            if (valid) {
                const double2 X = scale_point(v[u], sc, log_flag, eps, identity_scaler);
                if (scaled) scaled[sd.off + j] = X;
                cell = cell_of(X.y, cell_x0, cell_invw, C) * C + cell_of(X.x, cell_x0, cell_invw, C);
            }
            const unsigned act = __ballot_sync(0xffffffffu, valid);
            if (kRankMode == 2) {
                // sort (cell, lane) across the warp; invalid lanes carry the largest key and end up on top
                uint32_t key = valid ? ((uint32_t)cell << 5) | (uint32_t)lane : 0xffffffffu;
#pragma unroll
                for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, key, stride);
                        // ascending blocks where bit `size` of the lane is clear (the last merge is ascending everywhere)
                        const bool take_min = (((lane & stride) == 0) == (((lane & size) == 0) || size == 32));
                        key = take_min ? min(key, other) : max(key, other);
                    }
                }
                const uint32_t below = __shfl_up_sync(0xffffffffu, key, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || (below >> 5) != (key >> 5));
                const unsigned upto = heads & ((2u << lane) - 1u);              // heads at or below this lane (lane 31: all)
                const int seg = 31 - __clz(upto);                               // first lane of this lane's segment
                const unsigned above = heads & ~((2u << lane) - 1u);
                const int seg_end = above ? __ffs(above) - 1 : 32;
                const bool here = key != 0xffffffffu;                           // this sorted slot holds a real point
                const int scell = (int)(key >> 5);
                uint32_t before = 0u;
                if (here && seg == lane) { before = hw[scell]; hw[scell] = (uint16_t)(before + (uint32_t)(seg_end - lane)); }
                __syncwarp();
                before = __shfl_sync(0xffffffffu, before, seg);
                if (here && cellrank)
                    cellrank[sd.off + j0 + u * 32 + (int)(key & 31u)] = (uint32_t)scell | ((before + (uint32_t)(lane - seg)) << 16);
                continue;
            }
            uint32_t before = 0u;
            int leader = lane;
            unsigned m = 0u;
            if (kRankMode == 1) {
                m = act;
                for (int bit = 0; bit < cell_bits; ++bit) {
                    const unsigned v = __ballot_sync(0xffffffffu, (cell >> bit) & 1);
                    m &= ((cell >> bit) & 1) ? v : ~v;
                }
                if (!valid) m = 0u;
            } else if (valid) {
                m = __match_any_sync(act, cell);
            }
            if (valid) {
                leader = __ffs(m) - 1;
                if (lane == leader) { before = hw[cell]; hw[cell] = (uint16_t)(before + __popc(m)); }
            }
            __syncwarp();
            before = __shfl_sync(0xffffffffu, before, leader);
            // one 32-bit record per point: cell id (low half) and rank (high half; both fit 16 bits)
            if (valid && cellrank) cellrank[sd.off + j] = (uint32_t)cell | ((before + __popc(m & ((1u << lane) - 1u))) << 16);
        }
    }
    __syncthreads();
    const size_t row = (size_t)blockIdx.y * nb_max + blockIdx.x;
    uint32_t *bh = blockhist + row * ncells;
    uint16_t *wp = warpprefix ? warpprefix + row * kSortWarps * ncells : nullptr;
    for (int c = threadIdx.x; c < ncells; c += kSortThreads) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            if (wp) wp[w * ncells + c] = (uint16_t)t;      // same-cell points in earlier warps of this block
            t += s_hist[w * ncells + c];
        }
        bh[c] = t;
    }
}

// ------------------------------------------------------------------------------------------
// k_scan_narrow: grid (n_sets), 1024 threads, ncells <= 4096 (4 consecutive cells per thread).  Turns
// blockhist[set][b][c] into the first sorted position of block b's points of cell c (absolute), and writes
// cellstart[set][0..ncells].  Used for the 64x64-cell geometry, where a set has few sort blocks.
__global__ void __launch_bounds__(1024)
k_scan_narrow(const __grid_constant__ WaveDesc wd, int ncells, int nb_max,
              uint32_t *__restrict__ blockhist, uint32_t *__restrict__ cellstart,
              unsigned *__restrict__ counters) {
    const int set = blockIdx.x;
    if (set == 0 && threadIdx.x < kAllCounters) counters[threadIdx.x] = 0u;  // work-list (+ redo) counters of this wave
    const int nb = wd.set[set].nb;
    uint32_t *bh = blockhist + (size_t)set * nb_max * ncells;
    constexpr int CPT = 4;
    const int c0 = threadIdx.x * CPT;
    uint32_t tot[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) tot[k] = 0u;
    for (int b = 0; b < nb; ++b)
#pragma unroll
        for (int k = 0; k < CPT; ++k)
            if (c0 + k < ncells) tot[k] += bh[(size_t)b * ncells + c0 + k];
    uint32_t tsum = 0u;
#pragma unroll
    for (int k = 0; k < CPT; ++k) tsum += tot[k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __shared__ uint32_t s_w[32];
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t v = s_w[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_w[lane] = iv - v;
    }
    __syncthreads();
    uint32_t run = s_w[w] + inc - tsum;
    uint32_t *cs = cellstart + (size_t)set * (ncells + 1);
    uint32_t base[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        base[k] = run;
        if (c0 + k < ncells) cs[c0 + k] = run;
        run += tot[k];
    }
    if (threadIdx.x == 0) cs[ncells] = (uint32_t)wd.set[set].n;
    for (int b = 0; b < nb; ++b)
#pragma unroll
        for (int k = 0; k < CPT; ++k)
            if (c0 + k < ncells) {
                uint32_t t = bh[(size_t)b * ncells + c0 + k];
                bh[(size_t)b * ncells + c0 + k] = base[k];
                base[k] += t;
            }
}

// k_scan_wide: grid (ceil(ncells/1024), n_sets), 1024 threads, one cell per thread (128x128 cells, or a set
// with many sort blocks).  Turns blockhist[set][b][c]
// into the number of cell-c points in EARLIER blocks of the set (one coalesced read+write pass), leaves the
// cell totals in cellstart[set][c]; the last block of a set to finish (ticket) replaces those totals with
// their exclusive scan and writes cellstart[set][ncells] = n.  Sorted position of a point =
// cellstart[cell] + blockhist[block][cell] + warpprefix + rank (k_place).
__global__ void __launch_bounds__(1024)
k_scan_wide(const __grid_constant__ WaveDesc wd, int ncells, int nb_max,
            uint32_t *__restrict__ blockhist, uint32_t *__restrict__ cellstart,
            unsigned *__restrict__ counters, unsigned *__restrict__ ticket) {
    const int set = blockIdx.y;
    if (set == 0 && blockIdx.x == 0 && threadIdx.x < kAllCounters) counters[threadIdx.x] = 0u;  // work-list (+ redo) counters of this wave
    const int nb = wd.set[set].nb;
    uint32_t *cs = cellstart + (size_t)set * (ncells + 1);
    const int c = blockIdx.x * 1024 + threadIdx.x;
    if (c < ncells) {
        uint32_t *bh = blockhist + (size_t)set * nb_max * ncells + c;
        uint32_t run = 0u;
        int b = 0;
        for (; b + 4 <= nb; b += 4) {                      // four independent loads in flight
            uint32_t t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = bh[(size_t)(b + u) * ncells];
#pragma unroll
            for (int u = 0; u < 4; ++u) { bh[(size_t)(b + u) * ncells] = run; run += t[u]; }
        }
        for (; b < nb; ++b) { const uint32_t t = bh[(size_t)b * ncells]; bh[(size_t)b * ncells] = run; run += t; }
        cs[c] = run;
    }
    __shared__ bool s_last;
    __shared__ uint32_t s_w[32];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();               // cumulative: orders the block's cs[] stores before the ticket
        s_last = (atomicAdd(&ticket[set], 1u) == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    // exclusive scan of the ncells totals: thread t owns cpt consecutive cells
    constexpr int kCptMax = kMaxCells / 1024;
    const int cpt = (ncells + 1023) / 1024;
    const int c0 = threadIdx.x * cpt;
    uint32_t tot[kCptMax];
    uint32_t tsum = 0u;
#pragma unroll
    for (int k = 0; k < kCptMax; ++k) {
        tot[k] = (k < cpt && c0 + k < ncells) ? __ldcg(cs + c0 + k) : 0u;
        tsum += tot[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t v = s_w[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_w[lane] = iv - v;
    }
    __syncthreads();
    uint32_t run = s_w[w] + inc - tsum;
#pragma unroll
    for (int k = 0; k < kCptMax; ++k) {
        if (k < cpt && c0 + k < ncells) cs[c0 + k] = run;
        run += tot[k];
    }
    if (threadIdx.x == 0) { cs[ncells] = (uint32_t)wd.set[set].n; ticket[set] = 0u; }   // self-resetting
}

// 32-bit fixed-point copy of a scaled coordinate for the fp32 path.  Coordinates beyond [q_lo, q_hi] (2h outside the
// grid's extent: they cannot be within h of any node) are clamped so that differences to in-reach nodes never wrap.
__device__ __forceinline__ uint32_t quantise(double x, const GridGeom &g) {
    const double c = fmin(fmax(x, g.q_lo), g.q_hi);
    // rint(v) mod 2^32 without a 64-bit conversion: adding 1.5 * 2^52 leaves round-to-nearest(v) in the low
    // mantissa bits (|v| < 2^51 here: at most (extent/h + 4) * 2^29)
    return (uint32_t)__double2loint((c - g.gmin) * g.q_scale + 6755399441055744.0);
}

// ------------------------------------------------------------------------------------------
// k_place: grid (ceil(n_max/256), n_sets), 256 threads, one point per thread.  The point's stable sorted
// position is  cell start + earlier blocks (k_scan) + earlier warps of its block (k_scale_count) + its rank inside its
// warp sub-chunk (k_scale_count); the scaled coordinates are re-derived from the raw point with the same
// arithmetic as k_scale_count and written straight to that slot.  No shared memory, no ranking.
__global__ void __launch_bounds__(256)
k_place(const __grid_constant__ WaveDesc wd, const RawStat *__restrict__ rstat, int log_flag, double eps,
        int identity_scaler, int ncells, int chw, int nb_max, const uint32_t *__restrict__ cellrank,
        const uint16_t *__restrict__ warpprefix,
        const uint32_t *__restrict__ blockbase, const uint32_t *__restrict__ cellstart,
        double2 *__restrict__ sorted, uint2 *__restrict__ sorted_q, const GridGeom g) {
    __shared__ Scaler s_sc;                 // two fp64 divisions: once per block, not once per point
    const SetDesc sd = wd.set[blockIdx.y];
    if (threadIdx.x == 0) s_sc = make_scaler(rstat, sd, identity_scaler);
    __syncthreads();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;          // a set holds < 2^29 points (check_set)
    if (j >= (uint32_t)sd.n) return;
    const double2 v = reinterpret_cast<const double2 *>(wd.raw[sd.rs_self].p)[j];
    const uint32_t cr = cellrank[sd.off + j];
    const int cell = (int)(cr & 0xffffu);
    const uint32_t rank = cr >> 16;
    const uint32_t wchunk = j / (uint32_t)chw;            // global warp sub-chunk index of the point
    const size_t row = (size_t)blockIdx.y * nb_max + (size_t)(wchunk / kSortWarps);
    // cellstart is null when k_scan_narrow left absolute positions in blockbase
    const uint32_t pos = (cellstart ? cellstart[(size_t)blockIdx.y * (ncells + 1) + cell] : 0u) +
                         blockbase[row * ncells + cell] +
                         warpprefix[(row * kSortWarps + (size_t)(wchunk % kSortWarps)) * ncells + cell] + rank;
    const double2 X = scale_point(v, s_sc, log_flag, eps, identity_scaler);
    sorted[sd.off + pos] = X;
    if (sorted_q) sorted_q[sd.off + pos] = make_uint2(quantise(X.x, g), quantise(X.y, g));
}

// ------------------------------------------------------------------------------------------
// Per-node epilogue shared by the density kernels: density -> **gamma + eps.
// gamma_mode: 1 gamma==1, 2 gamma==0.5 (sqrt), 3 gamma==0.25 (sqrt of sqrt), 0 general pow().
__device__ __forceinline__ double gamma_eps(double dens, const GridGeom &g) {
    double z;
    if (g.gamma_mode == 1) z = dens;
    else if (g.gamma_mode == 2) z = sqrt(dens);
    else if (g.gamma_mode == 3) z = sqrt(sqrt(dens));
    else z = pow(dens, g.gamma);                                   // KDE_distance.py:71
    return z + g.eps;                                              // :75
}

// Cells a 4x4-node tile has to visit: every point within h of one of its nodes lies in cell rows
// [cy_lo, cy_hi] (cell_of is monotone; `margin` covers the rounding of the bounds), and inside cell row
// cy only in the columns tile_row_run() returns: a row whose y-interval is dy away from the tile's
// y-extent can only hold in-support points within sqrt(h^2 - dy^2) of its x-extent (a disc, not a box).
struct TileCells { int cy_lo, cy_hi; double gx0, xhi, gy0, yhi; };
__device__ __forceinline__ TileCells tile_cells(const GridGeom &g, int ix0, int iy0) {
    TileCells t;
    t.gx0 = (double)ix0 * g.step + g.gmin;
    t.gy0 = (double)iy0 * g.step + g.gmin;
    t.xhi = (double)min(ix0 + kTile - 1, g.R - 1) * g.step + g.gmin;
    t.yhi = (double)min(iy0 + kTile - 1, g.R - 1) * g.step + g.gmin;
    t.cy_lo = cell_of(t.gy0 - g.h - g.margin, g.cell_x0, g.cell_invw, g.C);
    t.cy_hi = cell_of(t.yhi + g.h + g.margin, g.cell_x0, g.cell_invw, g.C);
    return t;
}
// Candidates of cell row cy (tc.cy_lo <= cy <= tc.cy_hi): the cells [cx_lo, cx_hi] that can reach the tile.
// With g.use_moments the cells [mlo, mlo+mcnt) in the middle of that range -- those whose EVERY point lies
// within h of EVERY node of the tile -- are left out of the point runs: inside the support the Epanechnikov
// term is a polynomial, so such a cell enters through its moments (accumulate_cell) at the cost of one
// point.  What remains are two point runs (s1,l1) and (s2,l2) left and right of the moment cells.
// Bounds of sqrt(x), x >= 0, from the fp32 unit: the cell ranges only have to be conservative, and an fp64
// square root per cell row is a measurable part of the per-tile prologue (rel. error of the fp32 value <= 1.2e-7).
__device__ __forceinline__ double sqrt_up(double x) { return (double)(sqrtf((float)x) * 1.000001f); }
__device__ __forceinline__ double sqrt_down(double x) { return (double)(sqrtf((float)x) * 0.999999f); }
struct RowRuns { uint32_t s1, l1, s2, l2; int mlo, mcnt; };
__device__ __forceinline__ RowRuns tile_row_runs(const GridGeom &g, const TileCells &tc,
                                                 const uint32_t *__restrict__ cs, int cy, const bool moments) {
    // nominal y-interval of the row; the border rows also hold everything beyond the grid's extent, which
    // only lies farther from the tile than the nominal edge used here
    const double row_lo = g.cell_x0 + (double)cy * g.cell_w, row_hi = row_lo + g.cell_w;
    const double dy = fmax(fmax(row_lo - tc.yhi, tc.gy0 - row_hi) - g.margin, 0.0);
    const double rx2 = g.h * g.h - dy * dy;
    RowRuns rr = {0u, 0u, 0u, 0u, 0, 0};
    if (rx2 > 0.0) {
        const double rx = sqrt_up(rx2) + g.margin;
        const int cx_lo = cell_of(tc.gx0 - rx, g.cell_x0, g.cell_invw, g.C);
        const int cx_hi = cell_of(tc.xhi + rx, g.cell_x0, g.cell_invw, g.C);
        int mlo = cx_hi + 1, mhi = cx_hi;                    // no moment cells
        if (moments && cy > 0 && cy < g.C - 1) {             // border cells also hold out-of-extent points
            // farthest a point of this row can be from a node of the tile in y, then the x-reach left for it
            const double fy = fmax(tc.yhi - row_lo, row_hi - tc.gy0) + g.margin;
            const double fx2 = g.h * g.h - fy * fy;
            if (fx2 > 0.0) {
                const double fx = sqrt_down(fx2) - g.margin;
                // cell c = [x0 + c w, x0 + (c+1) w] qualifies iff  xhi - (x0 + c w) <= fx  and  x0 + (c+1) w - gx0 <= fx
                const int lo = max(__double2int_ru((tc.xhi - fx - g.cell_x0) * g.cell_invw + 1e-9), max(cx_lo, 1));
                const int hi = min(__double2int_rd((tc.gx0 + fx - g.cell_x0) * g.cell_invw - 1e-9) - 1, min(cx_hi, g.C - 2));
                if (lo <= hi) { mlo = lo; mhi = hi; }
            }
        }
        rr.s1 = cs[cy * g.C + cx_lo];
        rr.l1 = cs[cy * g.C + mlo] - rr.s1;
        rr.s2 = cs[cy * g.C + mhi + 1];
        rr.l2 = cs[cy * g.C + cx_hi + 1] - rr.s2;
        rr.mlo = mlo;
        rr.mcnt = mhi - mlo + 1;
    }
    return rr;
}

// Work-list layout.  Tiles are bucketed by candidate count so that the persistent density kernel can
// hand them out heaviest first (longest-processing-time order keeps the tail short):
//   class 0: >= 2^H candidates, H = heavy_log2 = 13 (one tile per block), class c in 1..9: [2^(H-c), 2^(H+1-c)) (per warp),
//   class 9 also takes everything below 16.
// counters: [group*10 + class] entries per list, then the block ticket and the warp ticket (zeroed by k_scan each wave).

__device__ __forceinline__ int tile_class(uint32_t total, int heavy_log2) {
    const int lg = 31 - __clz((int)total);          // floor(log2(total)), total >= 1
    return lg >= heavy_log2 ? 0 : min(heavy_log2 - lg, kNumClasses - 1);
}

// k_tile_classify: grid (ceil(tiles/32), n_sets), 256 threads = 32 tiles x 8 lanes.  The 8 lanes of a tile
// fetch its cell-row runs in parallel (one L2 round trip instead of a dependent chain) and add up the
// candidate count.  A tile with none gets its 16 node values written here (density 0), the others are
// appended to the work list of their weight class; the order inside a class is arbitrary and does not
// affect any result.
constexpr int kClassifyLanes = 8;
__global__ void __launch_bounds__(256)
k_tile_classify(const __grid_constant__ WaveDesc wd, const GridGeom g,
                const uint32_t *__restrict__ cellstart, double *__restrict__ dens,
                double *__restrict__ pz, double *__restrict__ tilesum, uint32_t *__restrict__ worklist,
                unsigned *__restrict__ counters, unsigned capacity,
                unsigned long long *__restrict__ work_counter, uint8_t *__restrict__ tileflag, const SplitBufs sb) {
    const int nty = g.tile_col_end - g.tile_col_begin;
    const int ntl = (g.tile_row_end - g.tile_row_begin) * nty;
    const int tl = blockIdx.x * (blockDim.x / kClassifyLanes) + threadIdx.x / kClassifyLanes;
    const int sub = threadIdx.x % kClassifyLanes;
    const int set = blockIdx.y;
    const bool live = tl < ntl;
    const int tx = g.tile_row_begin + (live ? tl / nty : 0), ty = g.tile_col_begin + (live ? tl % nty : 0);
    const int ix0 = tx * kTile, iy0 = ty * kTile;
    uint32_t total = 0u;
    if (live) {
        const TileCells tc = tile_cells(g, ix0, iy0);
        const uint32_t *__restrict__ cs = cellstart + (size_t)set * (g.C * g.C + 1);
        for (int cy = tc.cy_lo + sub; cy <= tc.cy_hi; cy += kClassifyLanes) {
            const RowRuns rr = tile_row_runs(g, tc, cs, cy, g.use_moments != 0);
            total += rr.l1 + rr.l2 + (uint32_t)rr.mcnt;
        }
    }
    total += __shfl_xor_sync(0xffffffffu, total, 1);
    total += __shfl_xor_sync(0xffffffffu, total, 2);
    total += __shfl_xor_sync(0xffffffffu, total, 4);
    // List slots are claimed per BLOCK: the 32 tiles of a block count their entries in shared memory and one thread
    // per class adds the block's total to the global counter.  One global atomic per tile had every non-empty tile
    // of a wave (150 000 at the headline size) queue up on the same ten addresses -- that, not the arithmetic, was
    // the kernel's time.  A block's tiles belong to one set, hence to one group: kNumClasses lists.
    __shared__ unsigned s_cnt[kNumClasses], s_base[kNumClasses];
    if (threadIdx.x < kNumClasses) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const int grp = set / g.sets_per_group;
    int cls = -1;
    uint32_t K = 1u, pbase = 0u, local = 0u;
    if (live) {
        const int tile = tx * g.NT + ty;
        if (tileflag && sub == 0) tileflag[(size_t)set * g.NT * g.NT + tile] = (total == 0u) ? 1 : 0;
        if (total == 0u) {
            const double P0 = gamma_eps(0.0, g);
#pragma unroll
            for (int k = sub; !tileflag && k < kTile * kTile; k += kClassifyLanes) {      // two nodes per lane
                const int ix = ix0 + k / kTile, iy = iy0 + k % kTile;
                if (ix < g.R && iy < g.R) {
                    const size_t n = (size_t)set * g.R * g.R + (size_t)ix * g.R + iy;
                    dens[n] = 0.0;
                    pz[n] = P0;
                }
            }
            if (sub == 0) {
                const int nv = min(kTile, g.R - ix0) * min(kTile, g.R - iy0);
                tilesum[(size_t)set * g.NT * g.NT + tile] = P0 * (double)nv;
            }
        } else if (sub == 0) {
            cls = tile_class(total, g.heavy_log2);
            const int split_log2 = g.split_cfg & 0xff, part_log2 = g.split_cfg >> 8;
            if (cls == 0 && total >= (1u << split_log2)) {
                K = min((uint32_t)kMaxParts, (total + (1u << part_log2) - 1u) >> part_log2);
                pbase = atomicAdd(&counters[kSplitParts], K);
                if (pbase + K > sb.parts_cap) K = 1u;                  // out of part slots: process it whole
            }
            local = atomicAdd(&s_cnt[cls], K);
        }
    }
    __syncthreads();
    if (threadIdx.x < kNumClasses && s_cnt[threadIdx.x] > 0u)
        s_base[threadIdx.x] = atomicAdd(&counters[grp * kNumClasses + threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (cls >= 0) {
        const int list = grp * kNumClasses + cls;
        const uint32_t entry = ((uint32_t)set << 24) | (uint32_t)(tx * g.NT + ty);
        const unsigned slot = s_base[cls] + local;
        for (uint32_t k = 0; k < K; ++k) {
            worklist[(size_t)list * capacity + slot + k] = entry;
            if (cls == 0) sb.partinfo[(size_t)grp * capacity + slot + k] = make_uint2(k | (K << 8), pbase);
        }
    }
    if (work_counter) {
        unsigned long long t = (live && sub == 0) ? total : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0 && t)
            atomicAdd(work_counter, t * (kTile * kTile));
    }
}

// k_tile_rowwork: grid (NT tile rows), 128 threads.  Estimated density work of one tile row over all sets
// of the wave: candidates plus a fixed per-tile cost.  Used to cut the node rows of one huge evaluation
// into equally heavy bands, one per rank (every rank computes the same numbers from the same sorted data).
__global__ void __launch_bounds__(128)
k_tile_rowwork(const __grid_constant__ WaveDesc wd, const GridGeom g,
               const uint32_t *__restrict__ cellstart, double *__restrict__ rowwork) {
    __shared__ double s_w[32];
    const int tx = blockIdx.x;
    double work = 0.0;
    for (int idx = threadIdx.x; idx < wd.n_sets * g.NT; idx += blockDim.x) {
        const int set = idx / g.NT, ty = idx % g.NT;
        const TileCells tc = tile_cells(g, tx * kTile, ty * kTile);
        const uint32_t *__restrict__ cs = cellstart + (size_t)set * (g.C * g.C + 1);
        uint32_t total = 0u;
        for (int cy = tc.cy_lo; cy <= tc.cy_hi; ++cy) {
            const RowRuns rr = tile_row_runs(g, tc, cs, cy, g.use_moments != 0);
            total += rr.l1 + rr.l2 + (uint32_t)rr.mcnt;
        }
        work += (double)total + (total ? 64.0 : 2.0);
    }
    work = block_sum(work, s_w);
    if (threadIdx.x == 0) rowwork[tx] = work;
}

// Clamp-and-accumulate for the Epanechnikov term t = 1 - d^2/h^2:  acc += max(t, 0).
// The clamp is ONE integer max on the high word (ALU pipe, otherwise idle here): hi' = max(hi32(t), 0).
// For t >= 0 nothing changes; for t < 0 the value collapses to {0, lo32(t)}, a subnormal below 2^-1042.
// Genuine terms are 0 or >= 2^-53 (t = 1 - u^2 with u^2 < 1), so that residue can never reach the last bit
// of a non-zero sum, and a sum made only of residues (< N * 2^-1042) is recognised in the epilogue and
// written as the exact 0 it stands for (kResidueFloor).  A pair costs two fp64-pipe instructions (DFMA for
// t, DADD into the accumulator) plus that one ALU op -- no compare/select chain, no mask registers.
constexpr double kResidueFloor = 1e-280;
__device__ __forceinline__ double clamp_nonneg(double t) {
    return __hiloint2double(max(__double2hiint(t), 0), __double2loint(t));
}
__device__ __forceinline__ double unresidue(double S) { return S < kResidueFloor ? 0.0 : S; }

constexpr int kRedStride = kDensThreads + 8;   // padded row of the block-level transposed reduction buffer
constexpr int kWarpRedStride = 33;             // padded row of the warp-level one
constexpr int kRedDoubles = kTile * kTile * kRedStride;   // >= 4 warps * 16 * 33

// One candidate point against the tile's 16 nodes (46 fp64-pipe + 16 FP32-pipe instructions).
__device__ __forceinline__ void accumulate_point(const double2 p, const double gx0, const double gy0,
                                                 const double inv_h, const double c1, const double c2,
                                                 const double c3, double (&acc)[kTile][kTile]) {
    const double ux0 = (p.x - gx0) * inv_h;      // node i sits at i*dh - ux0 (units of h)
    const double uy0 = (p.y - gy0) * inv_h;
    const double u1 = c1 - ux0, u2 = c2 - ux0, u3 = c3 - ux0;
    const double v1 = c1 - uy0, v2 = c2 - uy0, v3 = c3 - uy0;
    const double a[kTile] = {fma(-ux0, ux0, 1.0), fma(-u1, u1, 1.0), fma(-u2, u2, 1.0), fma(-u3, u3, 1.0)};
    const double w[kTile] = {-uy0, -v1, -v2, -v3};
    const double q[kTile] = {uy0, v1, v2, v3};
#pragma unroll
    for (int i = 0; i < kTile; ++i) {
        double t[kTile];
#pragma unroll
        for (int j = 0; j < kTile; ++j) t[j] = fma(w[j], q[j], a[i]);
#pragma unroll
        for (int j = 0; j < kTile; ++j) acc[i][j] += clamp_nonneg(t[j]);
    }
}

// One cell that lies wholly inside the support of all 16 nodes, through its moments about the cell centre
// (units of h): m = {n, Sx, Sy, Sxx, Syy} with dx_k = (x_k - cx)/h.  For node (i,j) at u_i = (gx_i - cx)/h,
// v_j = (gy_j - cy)/h:   sum_k [1 - (dx_k-u_i)^2 - (dy_k-v_j)^2] = A_i - B_j,
//   A_i = (n - Sxx) + u_i (2 Sx - n u_i),   B_j = Syy - v_j (2 Sy - n v_j).
constexpr int kMomentDoubles = 5;
__device__ __forceinline__ void accumulate_cell(const double *__restrict__ m, const double cx, const double cy,
                                                const double gx0, const double gy0, const double inv_h,
                                                const double c1, const double c2, const double c3,
                                                double (&acc)[kTile][kTile]) {
    const double n = m[0];
    if (n == 0.0) return;
    const double sx2 = 2.0 * m[1], sy2 = 2.0 * m[2], nmxx = n - m[3], syy = m[4];
    const double u0 = (gx0 - cx) * inv_h, v0 = (gy0 - cy) * inv_h;
    const double u[kTile] = {u0, u0 + c1, u0 + c2, u0 + c3};
    const double v[kTile] = {v0, v0 + c1, v0 + c2, v0 + c3};
    double A[kTile], B[kTile];
#pragma unroll
    for (int i = 0; i < kTile; ++i) {
        A[i] = fma(u[i], fma(-n, u[i], sx2), nmxx);
        B[i] = fma(-v[i], fma(-n, v[i], sy2), syy);
    }
#pragma unroll
    for (int i = 0; i < kTile; ++i)
#pragma unroll
        for (int j = 0; j < kTile; ++j) acc[i][j] += A[i] - B[j];
}
__device__ __forceinline__ double cell_centre(const GridGeom &g, int c) {
    return g.cell_x0 + ((double)c + 0.5) * g.cell_w;
}

// k_cell_moments: grid (ncells, n_sets), 256 threads, one block per cell: the five moments accumulate_cell()
// needs, summed thread-strided, folded with a fixed butterfly per warp and then warp after warp
// (bit-reproducible).  A block and not a warp per cell: real clouds put ~1 % of a 10 M-point set into its fullest
// cell, and one warp walking 100 k points alone was 0.6 ms -- a fixed cost that did not shrink when the
// evaluation was spread over more GPUs.
constexpr int kMomentThreads = 256;
__global__ void __launch_bounds__(kMomentThreads)
k_cell_moments(const __grid_constant__ WaveDesc wd, const GridGeom g, const double2 *__restrict__ sorted,
               const uint32_t *__restrict__ cellstart, double *__restrict__ moments) {
    __shared__ double s_part[kMomentThreads / 32][4];
    const int set = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncells = g.C * g.C;
    const int cell = blockIdx.x;
    const uint32_t *__restrict__ cs = cellstart + (size_t)set * (ncells + 1);
    const uint32_t b = cs[cell], e = cs[cell + 1];
    double *m = moments + ((size_t)set * ncells + cell) * kMomentDoubles;
    if (e == b) {                                          // block-uniform
        if (threadIdx.x < kMomentDoubles) m[threadIdx.x] = 0.0;
        return;
    }
    const double cx = cell_centre(g, cell % g.C), cy = cell_centre(g, cell / g.C);
    const double2 *__restrict__ pts = sorted + wd.set[set].off;
    double sx = 0.0, sy = 0.0, sxx = 0.0, syy = 0.0;
    for (uint32_t k = b + threadIdx.x; k < e; k += kMomentThreads) {
        const double2 p = pts[k];
        const double dx = (p.x - cx) * g.inv_h, dy = (p.y - cy) * g.inv_h;
        sx += dx; sy += dy;
        sxx = fma(dx, dx, sxx); syy = fma(dy, dy, syy);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sxx = warp_sum(sxx); syy = warp_sum(syy);
    if (lane == 0) { s_part[warp][0] = sx; s_part[warp][1] = sy; s_part[warp][2] = sxx; s_part[warp][3] = syy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < kMomentThreads / 32; ++w)
            for (int q = 0; q < 4; ++q) t[q] += s_part[w][q];
        m[0] = (double)(e - b); m[1] = t[0]; m[2] = t[1]; m[3] = t[2]; m[4] = t[3];
    }
}

// Work-list bookkeeping of one block: s_pre[l] = first item of list l in processing order (n_lists + 1 entries).
// Phase A walks the class-0 list of every group (group-major); phase B walks, group by group, classes 1..9.
// The list of an item is found by bisection, with no state carried across tiles (the candidate loops keep
// every register they had before the lists were grouped).
__device__ __forceinline__ uint32_t list_entry(const uint32_t *__restrict__ worklist, unsigned capacity,
                                               unsigned item, const unsigned *s_pre, const uint8_t *s_list,
                                               int n_lists) {
    int lo = 0, hi = n_lists - 1;                 // largest l with s_pre[l] <= item
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_pre[mid] <= item) lo = mid; else hi = mid - 1;
    }
    return worklist[(size_t)s_list[lo] * capacity + (item - s_pre[lo])];
}
// The same, also returning which list the item came from (the fp32 path re-queues a tile on the same list of the
// redo set when it cannot certify a node).
__device__ __forceinline__ uint32_t list_entry_ex(const uint32_t *__restrict__ worklist, unsigned capacity,
                                                  unsigned item, const unsigned *s_pre, const uint8_t *s_list,
                                                  int n_lists, int &list, unsigned &index) {
    int lo = 0, hi = n_lists - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_pre[mid] <= item) lo = mid; else hi = mid - 1;
    }
    list = s_list[lo];
    index = item - s_pre[lo];
    return worklist[(size_t)s_list[lo] * capacity + index];
}

// The share of one warp (of the four of a block) in cell-row run e of a tile, for the part that covers chunks
// [c0, c1) of the tile's flattened chunk sequence: the run holds chunks [cpre_e, cpre_e1); inside the part the
// chunks are dealt to the warps round-robin, (chunk - c0) mod 4 == warp.  Returns the first candidate of the warp's
// first chunk and how many candidates follow it up to the end of the run or of the part (0: nothing).
template <int CHUNK_LOG2 = 5>
__device__ __forceinline__ void part_run(uint32_t cpre_e, uint32_t cpre_e1, uint32_t rs_e, uint32_t len_e,
                                         uint32_t c0, uint32_t c1, int warp, uint32_t &start, uint32_t &rem) {
    const uint32_t lo = max(cpre_e, c0), hi = min(cpre_e1, c1);
    start = 0u;
    rem = 0u;
    if (lo < hi) {
        const uint32_t g0 = lo + (((uint32_t)warp - (lo - c0)) & 3u);
        if (g0 < hi) {
            const uint32_t off = (g0 - cpre_e) << CHUNK_LOG2;
            const uint32_t end = (hi == cpre_e1) ? len_e : (hi - cpre_e) << CHUNK_LOG2;     // a run's last chunk is ragged
            start = rs_e + off;
            rem = end - off;
        }
    }
}
// After a part's block-level reduction: `S` is valid in the owner thread of each node (owner == true for 16
// threads of the block).  K == 1: nothing to do.  Otherwise the partial goes to memory and the call returns true
// only in the block that finished the tile's last part, with S replaced by the sum of all parts in part order.
__device__ __forceinline__ bool combine_parts(const SplitBufs &sb, uint32_t k, uint32_t K, uint32_t pbase, bool owner,
                                              int node, double &S, unsigned *s_ticket) {
    if (K == 1u) return true;
    if (owner) sb.partial[(size_t)(pbase + k) * 16 + node] = S;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_ticket = atomicAdd(&sb.ticket[pbase], 1u);
    __syncthreads();
    if (*s_ticket != K - 1u) return false;
    __threadfence();
    if (owner) {
        const volatile double *vp = sb.partial + (size_t)pbase * 16 + node;
        double t = 0.0;
        for (uint32_t q = 0; q < K; ++q) t += vp[(size_t)q * 16];
        S = t;
    }
    if (threadIdx.x == 0) sb.ticket[pbase] = 0u;              // self-resetting
    return true;
}

// k_density_binned: persistent, (#SM x resident) blocks of 128 threads pulling 4x4-node tiles from
// the work list.  A tile's candidates are the sorted points of the cells that can reach it (one
// contiguous run per cell row; with MOMENTS two runs around the cells that enter through their moments).
// Lanes take candidates round-robin and keep the tile's 16 fp64 sums in registers:  t = 1 - ux^2 - uy^2 in
// units of h, added when non-negative (strict dist < h, sklearn:_binary_tree.pxi.tp:389-394; a zero term
// changes nothing).
//   phase A  the few very heavy tiles (class 0, >= 2^heavy_log2 candidates): one tile per BLOCK, the 32-candidate
//            chunks of its runs dealt to the four warps round-robin
//   phase B  everything else: one tile per WARP, no block-level synchronisation at all
// Both phases walk the work list in (set group, class) order -- see kMaxGroups.
// Each phase ends a tile with a fixed-shape reduction through shared memory, the 16 node values,
// their **gamma+eps and the tile's share of the normaliser.  Which warp or block processes a tile is
// decided at run time (atomic tickets) but never changes a result.
template <bool MOMENTS>
__global__ void __launch_bounds__(kDensThreads, 6)
k_density_binned(const __grid_constant__ WaveDesc wd, const GridGeom g,
                 const double2 *__restrict__ sorted, const uint32_t *__restrict__ cellstart,
                 const double *__restrict__ moments,
                 double *__restrict__ dens, double *__restrict__ pz, double *__restrict__ tilesum,
                 const uint32_t *__restrict__ worklist, unsigned *__restrict__ counters,
                 unsigned capacity, const SplitBufs sb) {
    __shared__ uint32_t s_rs[2 * kMaxRows];          // two point runs per cell row (left / right of the moment cells)
    __shared__ uint32_t s_len[2 * kMaxRows];         // their lengths
    __shared__ int s_mlo[kMaxRows], s_mcnt[kMaxRows], s_rows, s_cy0;
    __shared__ uint32_t s_cpre[2 * kMaxRows + 1];    // 32-candidate chunks before each run
    __shared__ double s_red[kRedDoubles];
    __shared__ unsigned s_item, s_ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncells = g.C * g.C;
    // processing order: phase A = class 0 of every group; phase B = for each group, classes 1..9
    __shared__ unsigned s_preA[kMaxGroups + 1], s_preB[kMaxGroups * (kNumClasses - 1) + 1];
    __shared__ uint8_t s_listA[kMaxGroups], s_listB[kMaxGroups * (kNumClasses - 1)];
    __shared__ unsigned s_cnt[kNumLists];
    if (threadIdx.x < kNumLists) s_cnt[threadIdx.x] = counters[threadIdx.x];     // one round trip for all lists
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned a = 0u, b = 0u;
        int nb_lists = 0;
#pragma unroll 1
        for (int grp = 0; grp < kMaxGroups; ++grp) {
            s_preA[grp] = a; s_listA[grp] = (uint8_t)(grp * kNumClasses);
            a += s_cnt[grp * kNumClasses];
#pragma unroll 1
            for (int c = 1; c < kNumClasses; ++c) {
                s_preB[nb_lists] = b; s_listB[nb_lists] = (uint8_t)(grp * kNumClasses + c);
                b += s_cnt[grp * kNumClasses + c];
                ++nb_lists;
            }
        }
        s_preA[kMaxGroups] = a;
        s_preB[nb_lists] = b;
    }
    __syncthreads();
    const unsigned n0 = s_preA[kMaxGroups];
    const unsigned nB = s_preB[kMaxGroups * (kNumClasses - 1)];
    const double c1 = g.dh, c2 = 2.0 * g.dh, c3 = 3.0 * g.dh;
    const size_t G = (size_t)g.R * g.R;

    // ---------------- phase A: block per tile ----------------
    for (;;) {
        if (n0 == 0u) break;
        if (threadIdx.x == 0) s_item = atomicAdd(&counters[kTicketA], 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= n0) break;
        int list;
        unsigned index;
        const uint32_t entry = list_entry_ex(worklist, capacity, item, s_preA, s_listA, kMaxGroups, list, index);
        const uint2 pinfo = sb.partinfo[(size_t)(list / kNumClasses) * capacity + index];
        const uint32_t part_k = pinfo.x & 0xffu, part_K = pinfo.x >> 8, pbase = pinfo.y;
        const int set = (int)(entry >> 24), tile = (int)(entry & 0xffffffu);
        const SetDesc sd = wd.set[set];
        const int tx = tile / g.NT, ty = tile % g.NT;
        const int ix0 = tx * kTile, iy0 = ty * kTile;
        const double gx0 = (double)ix0 * g.step + g.gmin, gy0 = (double)iy0 * g.step + g.gmin;
        if (warp == 0) {
            const uint32_t *__restrict__ cs = cellstart + (size_t)set * (ncells + 1);
            const TileCells tc = tile_cells(g, ix0, iy0);
            const int rows = tc.cy_hi - tc.cy_lo + 1;       // <= kMaxRows by the host's choice of C
            RowRuns rr = {0u, 0u, 0u, 0u, 0, 0};
            if (lane < rows) rr = tile_row_runs(g, tc, cs, tc.cy_lo + lane, MOMENTS);
            s_rs[2 * lane] = rr.s1;
            s_rs[2 * lane + 1] = rr.s2;
            s_len[2 * lane] = rr.l1;
            s_len[2 * lane + 1] = rr.l2;
            // 32-candidate chunks are dealt to the four warps round-robin ACROSS runs: s_cpre = chunks before a run
            const uint32_t c1n = (rr.l1 + 31u) >> 5, c2n = (rr.l2 + 31u) >> 5;
            uint32_t cinc = c1n + c2n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, cinc, o);
                if (lane >= o) cinc += t;
            }
            s_cpre[2 * lane] = cinc - c1n - c2n;
            s_cpre[2 * lane + 1] = cinc - c2n;
            if (lane == 31) s_cpre[2 * kMaxRows] = cinc;
            s_mlo[lane] = rr.mlo;
            s_mcnt[lane] = rr.mcnt;
            if (lane == 0) { s_rows = rows; s_cy0 = tc.cy_lo; }
        }
        __syncthreads();
        double acc[kTile][kTile];
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) acc[i][j] = 0.0;
        const double2 *__restrict__ pts = sorted + sd.off;
        // this part's chunks of the tile's flattened chunk sequence (a tile that is not split is one part)
        const uint32_t gtot = s_cpre[2 * kMaxRows];
        const uint32_t pc0 = (uint32_t)((uint64_t)gtot * part_k / part_K), pc1 = (uint32_t)((uint64_t)gtot * (part_k + 1u) / part_K);
        // run by run, each warp strides through every fourth 32-candidate chunk: the same loop as phase B (no
        // per-candidate run lookup), and only the last chunk of a run is ragged
        const int nruns = (MOMENTS ? 2 : 1) * s_rows;
        for (int k = 0; k < nruns; ++k) {
            const int e = MOMENTS ? k : 2 * k;               // without moment cells only the even entries are used
            uint32_t start, len;
            part_run(s_cpre[e], s_cpre[e + 1], s_rs[e], s_len[e], pc0, pc1, warp, start, len);
            const double2 *__restrict__ rp = pts + start;
            uint32_t j = lane;
            double2 pn = make_double2(0.0, 0.0);
            if (j < len) pn = rp[j];
            while (j < len) {
                const double2 p = pn;
                j += kDensThreads;
                if (j < len) pn = rp[j];                     // the lane's next candidate is in flight
                accumulate_point(p, gx0, gy0, g.inv_h, c1, c2, c3, acc);
            }
        }
        if (MOMENTS && part_k == 0u) {                       // the moment cells go with the first part
            const double *__restrict__ mom = moments + (size_t)set * ncells * kMomentDoubles;
            const int rows = s_rows, cy0 = s_cy0;
            for (int rr = 0; rr < rows; ++rr) {
                const int mc = s_mcnt[rr], cy = cy0 + rr;
                for (int k = threadIdx.x; k < mc; k += kDensThreads) {
                    const int cx = s_mlo[rr] + k;
                    accumulate_cell(mom + (size_t)(cy * g.C + cx) * kMomentDoubles, cell_centre(g, cx),
                                    cell_centre(g, cy), gx0, gy0, g.inv_h, c1, c2, c3, acc);
                }
            }
        }
        // transpose through shared memory: 8 partial sums of 16 lanes per node, then a 3-step butterfly
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) s_red[(i * kTile + j) * kRedStride + threadIdx.x] = acc[i][j];
        __syncthreads();
        const int node = threadIdx.x >> 3, part = threadIdx.x & 7;
        double S = 0.0;
#pragma unroll
        for (int k = 0; k < kDensThreads / 8; ++k) S += s_red[node * kRedStride + part + 8 * k];
        S += __shfl_xor_sync(0xffffffffu, S, 1);
        S += __shfl_xor_sync(0xffffffffu, S, 2);
        S += __shfl_xor_sync(0xffffffffu, S, 4);
        if (!combine_parts(sb, part_k, part_K, pbase, part == 0, node, S, &s_ticket)) continue;   // not the last part
        const int ix = ix0 + node / kTile, iy = iy0 + node % kTile;
        double P = 0.0;
        if (part == 0 && ix < g.R && iy < g.R) {
            const double d = unresidue(S) * sd.knorm_over_n;   // exp(logsum + log_knorm - log N)
            P = gamma_eps(d, g);
            dens[(size_t)set * G + (size_t)ix * g.R + iy] = d;
            pz[(size_t)set * G + (size_t)ix * g.R + iy] = P;
        }
        P += __shfl_xor_sync(0xffffffffu, P, 8);       // nodes 4w..4w+3 live in lanes 0,8,16,24 of warp w
        P += __shfl_xor_sync(0xffffffffu, P, 16);
        __syncthreads();                               // everyone is done reading s_red
        if (lane == 0) s_red[warp] = P;
        __syncthreads();
        if (threadIdx.x == 0)
            tilesum[(size_t)set * g.NT * g.NT + tile] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
        __syncthreads();                               // s_item, s_rs, s_len, s_red free for the next tile
    }

    // ---------------- phase B: warp per tile ----------------
    double *__restrict__ wred = s_red + warp * (kTile * kTile * kWarpRedStride);
    __syncthreads();                                   // phase A's last readers of s_red are done
    for (;;) {
        unsigned item = 0u;
        if (lane == 0) item = atomicAdd(&counters[kTicketB], 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nB) break;
        const uint32_t entry = list_entry(worklist, capacity, item, s_preB, s_listB, kMaxGroups * (kNumClasses - 1));
        const int set = (int)(entry >> 24), tile = (int)(entry & 0xffffffu);
        const double knorm_over_n = wd.set[set].knorm_over_n;
        const double2 *__restrict__ pts = sorted + wd.set[set].off;
        const int tx = tile / g.NT, ty = tile % g.NT;
        const int ix0 = tx * kTile, iy0 = ty * kTile;
        const double gx0 = (double)ix0 * g.step + g.gmin, gy0 = (double)iy0 * g.step + g.gmin;
        const uint32_t *__restrict__ cs = cellstart + (size_t)set * (ncells + 1);
        const TileCells tc = tile_cells(g, ix0, iy0);
        const int rows = tc.cy_hi - tc.cy_lo + 1;
        RowRuns rr = {0u, 0u, 0u, 0u, 0, 0};
        if (lane < rows) rr = tile_row_runs(g, tc, cs, tc.cy_lo + lane, MOMENTS);
        double acc[kTile][kTile];
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) acc[i][j] = 0.0;
        constexpr int nparts = MOMENTS ? 2 : 1;        // without moment cells the second run is always empty
        for (int r = 0; r < rows; ++r) {
#pragma unroll 1
            for (int part = 0; part < nparts; ++part) {
                const uint32_t len = __shfl_sync(0xffffffffu, part ? rr.l2 : rr.l1, r);
                const double2 *__restrict__ rp = pts + __shfl_sync(0xffffffffu, part ? rr.s2 : rr.s1, r);
                uint32_t j = lane;
                double2 pn = make_double2(0.0, 0.0);
                if (j < len) pn = rp[j];
                while (j < len) {
                    const double2 p = pn;
                    j += 32;
                    if (j < len) pn = rp[j];           // the lane's next candidate is in flight
                    accumulate_point(p, gx0, gy0, g.inv_h, c1, c2, c3, acc);
                }
            }
            if (MOMENTS) {
                const int mc = __shfl_sync(0xffffffffu, rr.mcnt, r);
                const int mlo = __shfl_sync(0xffffffffu, rr.mlo, r);
                const int cy = tc.cy_lo + r;
                const double *__restrict__ mom = moments + (size_t)set * ncells * kMomentDoubles;
                for (int k = lane; k < mc; k += 32)
                    accumulate_cell(mom + (size_t)(cy * g.C + mlo + k) * kMomentDoubles, cell_centre(g, mlo + k),
                                    cell_centre(g, cy), gx0, gy0, g.inv_h, c1, c2, c3, acc);
            }
        }
        // warp-level transpose: lane l sums 16 lanes' partials of node l/2, then one butterfly step
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) wred[(i * kTile + j) * kWarpRedStride + lane] = acc[i][j];
        __syncwarp();
        const int node = lane >> 1, half = lane & 1;
        double S = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) S += wred[node * kWarpRedStride + half + 2 * k];
        S += __shfl_xor_sync(0xffffffffu, S, 1);
        const int ix = ix0 + node / kTile, iy = iy0 + node % kTile;
        double P = 0.0;
        if (half == 0 && ix < g.R && iy < g.R) {
            const double d = unresidue(S) * knorm_over_n;
            P = gamma_eps(d, g);
            dens[(size_t)set * G + (size_t)ix * g.R + iy] = d;
            pz[(size_t)set * G + (size_t)ix * g.R + iy] = P;
        }
#pragma unroll
        for (int o = 2; o < 32; o <<= 1) P += __shfl_xor_sync(0xffffffffu, P, o);   // 16 node values, fixed tree
        if (lane == 0) tilesum[(size_t)set * g.NT * g.NT + tile] = P;
    }
}


// ------------------------------------------------------------------------------------------
// k_density_binned32: the culled sum on the FP32 pipe.  Same work list, same cell runs, same tile shape as
// k_density_binned; the candidates come from the 32-bit fixed-point copy of the sorted points (k_place), 8 bytes
// each.  Per candidate and lane: two integer subtractions against the tile's first node (exact), two I2FP, eight
// FFMA/FMUL for the node offsets in units of h, four FFMA for a_i = 1 - ux_i^2, then per node
//     t = sat(a_i - uy_j^2)        FFMA.SAT: multiply, add and an exact ReLU in one instruction (t <= 1 always)
//     acc_ij += t                  FADD2 (packed f32x2: one issue slot for two nodes)
// i.e. 44 FMA-pipe cycles and ~50 issue slots where the fp64 loop takes 135 cycles (measured, kdejs_loop_peak).
//
// The loop is half as long as the fp64 one, so it can no longer hide an L2 round trip behind ONE trip of its own
// arithmetic: a tile's candidate runs are therefore flattened into one stream of 32-candidate trips that a small
// register queue fetches kPrefetch trips ahead (the bookkeeping of run boundaries lives on the fetch side, is
// warp-uniform and costs a few instructions per trip; the arithmetic side just consumes the queue).
//
// Accumulation: fp32 partial sums are folded every kFoldTrips trips into a second fp32 level (and, for the heavy
// tiles a whole block works on, every kFoldTrips^2 trips into a third), so no fp32 sum ever runs over more than
// 16 terms of its own level; the cross-lane reduction is fp64 and fixed-shape (bit-reproducible like the fp64 path).
//
// Exact zeros stay exact zeros.  The error of a computed t is below kBias32/2 (coordinates: 2^-29 h quantisation,
// one I2FP rounding of |u| <= 2.5, fp32 node offsets), so
//   * a node whose sum reaches  thr = kBias32 * max(candidates, 1024)  certainly has an in-support point;
//   * otherwise the tile is walked a second time with every a_i raised by kBias32: a node whose biased sum is still
//     exactly 0 certainly has none (every true t <= -kBias32/2) and is written as the exact 0;
//   * a node that is neither (0 < true |t| of its nearest candidate < ~1e-6: a handful per evaluation) sends its
//     tile to the redo list, which k_density_binned (fp64) processes right after this kernel.
// Phase A (heavy tiles, block per tile) skips the biased walk and re-queues directly; it practically never happens.
constexpr int kFoldTrips = 16;
constexpr int kPrefetch = 3;                          // trips in flight per warp
constexpr float kBias32 = 9.5367431640625e-07f;       // 2^-20
constexpr uint32_t kFarQ = 0x40000000u;               // a padding candidate 2^30 units (>= 2 h, q_shift <= 29) LEFT of the
                                                      // tile's first node: out of every node's support, adds exactly 0

// Where the fp32 kernel takes its candidates from.  0 (default): a 32-bit fixed-point copy of the sorted points
// written by k_place (8 bytes per candidate to load; k_place scatters 24 instead of 16 bytes per point: 1.9 -> 2.8 us
// per evaluation).  1: the fp64 sorted points themselves, re-centred on the tile's first node in fp64 (2 DADD + 2
// DMUL on the otherwise idle fp64 pipe) and rounded to fp32 (2 F2F) -- no second copy, but measured slower overall:
// k_place -0.9 us, density +1.9 us per evaluation (16-byte candidates, a wider register queue, spills).
#ifndef KDEJS_FP32_FROM_DOUBLES
#define KDEJS_FP32_FROM_DOUBLES 0
#endif
constexpr bool kFp32FromDoubles = KDEJS_FP32_FROM_DOUBLES != 0;
struct TileOrigin32 {                // per tile: its first node in the candidates' representation
    uint32_t gxq, gyq;               // fixed point
    double gx0, gy0, inv_h;          // fp64
    float unit;                      // 2^-q_shift
};
#if KDEJS_FP32_FROM_DOUBLES
using Cand32 = double2;
__device__ __forceinline__ Cand32 far_candidate(const TileOrigin32 &o, const GridGeom &g) {
    return make_double2(o.gx0 - 2.0 * g.h, o.gy0);         // 2 h left of the first node: adds exactly 0 to every node
}
__device__ __forceinline__ void candidate_offsets(const Cand32 q, const TileOrigin32 &o, float &u0, float &v0) {
    u0 = (float)((q.x - o.gx0) * o.inv_h);                 // offsets from the tile's first node, units of h
    v0 = (float)((q.y - o.gy0) * o.inv_h);
}
#else
using Cand32 = uint2;
__device__ __forceinline__ Cand32 far_candidate(const TileOrigin32 &o, const GridGeom &) {
    return make_uint2(o.gxq - kFarQ, o.gyq);
}
__device__ __forceinline__ void candidate_offsets(const Cand32 q, const TileOrigin32 &o, float &u0, float &v0) {
    u0 = (float)(int)(q.x - o.gxq) * o.unit;
    v0 = (float)(int)(q.y - o.gyq) * o.unit;
}
#endif

__device__ __forceinline__ void accumulate_point32(const Cand32 q, const TileOrigin32 &o,
                                                   const float c1, const float c2, const float c3,
                                                   const float one, float (&acc)[kTile][kTile]) {
    float u0, v0;
    candidate_offsets(q, o, u0, v0);
    const float u1 = u0 - c1, u2 = u0 - c2, u3 = u0 - c3;
    const float v1 = v0 - c1, v2 = v0 - c2, v3 = v0 - c3;
    const float a[kTile] = {fmaf(-u0, u0, one), fmaf(-u1, u1, one), fmaf(-u2, u2, one), fmaf(-u3, u3, one)};
    const float v[kTile] = {v0, v1, v2, v3};
#pragma unroll
    for (int i = 0; i < kTile; ++i)
#pragma unroll
        for (int j = 0; j < kTile; j += 2) {
            const float2 t = make_float2(__saturatef(fmaf(-v[j], v[j], a[i])),
                                         __saturatef(fmaf(-v[j + 1], v[j + 1], a[i])));
            const float2 r = __fadd2_rn(make_float2(acc[i][j], acc[i][j + 1]), t);
            acc[i][j] = r.x;
            acc[i][j + 1] = r.y;
        }
}
// hi += lo; lo = 0   (packed adds)
__device__ __forceinline__ void fold32(float (&lo)[kTile][kTile], float (&hi)[kTile][kTile]) {
#pragma unroll
    for (int i = 0; i < kTile; ++i)
#pragma unroll
        for (int j = 0; j < kTile; j += 2) {
            const float2 r = __fadd2_rn(make_float2(hi[i][j], hi[i][j + 1]), make_float2(lo[i][j], lo[i][j + 1]));
            hi[i][j] = r.x; hi[i][j + 1] = r.y;
            lo[i][j] = 0.f; lo[i][j + 1] = 0.f;
        }
}
__device__ __forceinline__ void flush32(float (&a32)[kTile][kTile], double (&a64)[kTile][kTile]) {
#pragma unroll
    for (int i = 0; i < kTile; ++i)
#pragma unroll
        for (int j = 0; j < kTile; ++j) { a64[i][j] += (double)a32[i][j]; a32[i][j] = 0.f; }
}

// One walk of a warp over its flattened runs of a tile.  wstart / wlen: the warp's non-empty runs in processing
// order (per-warp shared memory): first candidate of the warp's first chunk in the run, and how many candidates the
// run holds from there on.  PAIR = false (phase A): a trip is 32 candidates, one per lane, and the warp takes every
// (STRIDE/32)-th chunk (the four warps of a block interleave).  PAIR = true (phase B, STRIDE 64): a trip is 64
// candidates, TWO consecutive ones per lane -- the fetch bookkeeping and the loop overhead are paid once per 64
// candidates (about a third of the instructions of a 32-candidate trip were not arithmetic).
// Returns the lane's 16 partial sums in `tot` (fp64).
struct CandPair { Cand32 a, b; };
template <int STRIDE, bool THREE_LEVEL, bool PAIR>
__device__ __forceinline__ void walk_tile32(const Cand32 *__restrict__ pts, const uint32_t *wstart, const uint32_t *wlen,
                                            const int nruns, const uint32_t total_trips, const int lane,
                                            const TileOrigin32 &o, const Cand32 far, const float f1,
                                            const float f2, const float f3, const float one,
                                            double (&tot)[kTile][kTile]) {
    constexpr int kPer = PAIR ? 2 : 1;            // candidates per lane and trip
    float lo[kTile][kTile], hi[kTile][kTile], top[kTile][kTile];
#pragma unroll
    for (int i = 0; i < kTile; ++i)
#pragma unroll
        for (int j = 0; j < kTile; ++j) { lo[i][j] = 0.f; hi[i][j] = 0.f; top[i][j] = 0.f; }
    // fetch side: a running pointer into the current run, the lane's remaining candidates in it and the run's
    // remaining trips; run boundaries cost a few warp-uniform instructions, a trip five
    int r_pf = 0;
    int rem = -0x40000000;                        // candidates of the current run at or after this lane's next one
    uint32_t trips_left = 0xffffffffu;            // (past the last run: nothing left to load, never switches again)
    const Cand32 *nxt = pts;
    if (nruns > 0) { rem = (int)wlen[0] - kPer * lane; nxt = pts + wstart[0] + kPer * lane; trips_left = (wlen[0] + STRIDE - 1u) / STRIDE; }
    auto fetch = [&]() -> CandPair {
        CandPair v{far, far};
        if (rem > 0) v.a = nxt[0];
        if (PAIR && rem > 1) v.b = nxt[1];
        nxt += STRIDE;
        rem -= STRIDE;
        if (--trips_left == 0u) {                             // warp-uniform, once per run
            ++r_pf;
            if (r_pf < nruns) {
                const uint32_t len = wlen[r_pf];
                rem = (int)len - kPer * lane;
                nxt = pts + wstart[r_pf] + kPer * lane;
                trips_left = (len + STRIDE - 1u) / STRIDE;
            } else {
                rem = -0x40000000;
                trips_left = 0xffffffffu;
            }
        }
        return v;
    };
    auto use = [&](const CandPair &c) {
        accumulate_point32(c.a, o, f1, f2, f3, one, lo);
        if (PAIR) accumulate_point32(c.b, o, f1, f2, f3, one, lo);
    };
    // kQ trips in flight (96 / 128 candidates per warp); the loop is unrolled by kQ so the queue is renamed, not
    // moved; fp32 partial sums are folded every kBlock trips (a level never sums more than 16 terms of its own): no
    // per-trip counter in the hot loop
    constexpr int kQ = PAIR ? 2 : 3, kBlock = PAIR ? 8 : 15;
    static_assert(kFoldTrips == 16 && kBlock * kPer <= kFoldTrips && kBlock % kQ == 0, "fold blocks");
    CandPair q[kQ];
#pragma unroll
    for (int i = 0; i < kQ; ++i) q[i] = fetch();
    uint32_t t = 0;
    int folds = kFoldTrips;
    while (t + kBlock <= total_trips) {
#pragma unroll 1
        for (int k = 0; k < kBlock / kQ; ++k) {
#pragma unroll
            for (int i = 0; i < kQ; ++i) { use(q[i]); q[i] = fetch(); }
        }
        t += kBlock;
        fold32(lo, hi);
        if (THREE_LEVEL && --folds == 0) { fold32(hi, top); folds = kFoldTrips; }
    }
    for (; t + kQ <= total_trips; t += kQ) {                   // fewer than kBlock trips left
#pragma unroll
        for (int i = 0; i < kQ; ++i) { use(q[i]); q[i] = fetch(); }
    }
#pragma unroll
    for (int i = 0; i < kQ - 1; ++i)
        if (t + i < total_trips) use(q[i]);
    fold32(lo, hi);
    if (THREE_LEVEL) fold32(hi, top);
#pragma unroll
    for (int i = 0; i < kTile; ++i)
#pragma unroll
        for (int j = 0; j < kTile; ++j) tot[i][j] = (double)(THREE_LEVEL ? top[i][j] : hi[i][j]);
}

// Phase A of the fp32 kernel (block per heavy tile): candidates per lane and trip, and with that the chunk the four
// warps deal out round-robin (KDEJS_PAIR_A=0: one candidate per lane, 32-candidate chunks).
#ifndef KDEJS_PAIR_A
#define KDEJS_PAIR_A 1
#endif
constexpr bool kPairA = KDEJS_PAIR_A != 0;
constexpr int kChunkALog2 = kPairA ? 6 : 5;
constexpr uint32_t kChunkA = 1u << kChunkALog2;
constexpr int kStrideA = (kDensThreads / 32) * (int)kChunkA;      // candidates between a warp's consecutive chunks
template <bool MOMENTS>
__global__ void __launch_bounds__(kDensThreads, 6)
k_density_binned32(const __grid_constant__ WaveDesc wd, const GridGeom g,
                   const Cand32 *__restrict__ sorted_q, const uint32_t *__restrict__ cellstart,
                   const double *__restrict__ moments,
                   double *__restrict__ dens, double *__restrict__ pz, double *__restrict__ tilesum,
                   const uint32_t *__restrict__ worklist, unsigned *__restrict__ counters,
                   unsigned capacity, uint32_t *__restrict__ redo_list, const float min_thr,
                   unsigned long long *__restrict__ stats, const SplitBufs sb, uint2 *__restrict__ redo_partinfo) {
    // stats (profiling only): [0] tiles, [1] tiles walked a second time, [2] warp trips of first walks,
    // [3] warp trips of second walks, [4] tiles re-queued for fp64
    __shared__ uint32_t s_rs[2 * kMaxRows];          // phase A: the tile's runs; phase B: per-warp compacted runs
    __shared__ uint32_t s_len[2 * kMaxRows];
    __shared__ uint32_t s_cpre[2 * kMaxRows + 1];    // phase A: 32-candidate chunks before each run
    __shared__ uint32_t s_wstart[kDensThreads / 32][2 * kMaxRows], s_wlen[kDensThreads / 32][2 * kMaxRows];
    __shared__ int s_mlo[kMaxRows], s_mcnt[kMaxRows], s_rows, s_cy0;
    __shared__ double s_red[kRedDoubles];
    __shared__ unsigned s_item, s_ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncells = g.C * g.C;
    __shared__ unsigned s_preA[kMaxGroups + 1], s_preB[kMaxGroups * (kNumClasses - 1) + 1];
    __shared__ uint8_t s_listA[kMaxGroups], s_listB[kMaxGroups * (kNumClasses - 1)];
    __shared__ unsigned s_cnt[kNumLists];
    if (threadIdx.x < kNumLists) s_cnt[threadIdx.x] = counters[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned a = 0u, b = 0u;
        int nb_lists = 0;
#pragma unroll 1
        for (int grp = 0; grp < kMaxGroups; ++grp) {
            s_preA[grp] = a; s_listA[grp] = (uint8_t)(grp * kNumClasses);
            a += s_cnt[grp * kNumClasses];
#pragma unroll 1
            for (int c = 1; c < kNumClasses; ++c) {
                s_preB[nb_lists] = b; s_listB[nb_lists] = (uint8_t)(grp * kNumClasses + c);
                b += s_cnt[grp * kNumClasses + c];
                ++nb_lists;
            }
        }
        s_preA[kMaxGroups] = a;
        s_preB[nb_lists] = b;
    }
    __syncthreads();
    const unsigned n0 = s_preA[kMaxGroups];
    const unsigned nB = s_preB[kMaxGroups * (kNumClasses - 1)];
    const double c1 = g.dh, c2 = 2.0 * g.dh, c3 = 3.0 * g.dh;
    const float f1 = g.q_dh, f2 = 2.0f * g.q_dh, f3 = 3.0f * g.q_dh;
    const float unit = g.q_unit;
    const size_t G = (size_t)g.R * g.R;
    unsigned *__restrict__ redo_counters = counters + kRedoBase;

    // ---------------- phase A: block per tile ----------------
    for (;;) {
        if (n0 == 0u) break;
        if (threadIdx.x == 0) s_item = atomicAdd(&counters[kTicketA], 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= n0) break;
        int list;
        unsigned index;
        const uint32_t entry = list_entry_ex(worklist, capacity, item, s_preA, s_listA, kMaxGroups, list, index);
        const uint2 pinfo = sb.partinfo[(size_t)(list / kNumClasses) * capacity + index];
        const uint32_t part_k = pinfo.x & 0xffu, part_K = pinfo.x >> 8, pbase = pinfo.y;
        const int set = (int)(entry >> 24), tile = (int)(entry & 0xffffffu);
        const SetDesc sd = wd.set[set];
        const int tx = tile / g.NT, ty = tile % g.NT;
        const int ix0 = tx * kTile, iy0 = ty * kTile;
        const double gx0 = (double)ix0 * g.step + g.gmin, gy0 = (double)iy0 * g.step + g.gmin;
        const TileOrigin32 org{kFp32FromDoubles ? 0u : quantise(gx0, g), kFp32FromDoubles ? 0u : quantise(gy0, g), gx0, gy0, g.inv_h, unit};
        if (warp == 0) {
            const uint32_t *__restrict__ cs = cellstart + (size_t)set * (ncells + 1);
            const TileCells tc = tile_cells(g, ix0, iy0);
            const int rows = tc.cy_hi - tc.cy_lo + 1;
            RowRuns rr = {0u, 0u, 0u, 0u, 0, 0};
            if (lane < rows) rr = tile_row_runs(g, tc, cs, tc.cy_lo + lane, MOMENTS);
            s_rs[2 * lane] = rr.s1;
            s_rs[2 * lane + 1] = rr.s2;
            s_len[2 * lane] = rr.l1;
            s_len[2 * lane + 1] = rr.l2;
            const uint32_t c1n = (rr.l1 + kChunkA - 1u) >> kChunkALog2, c2n = (rr.l2 + kChunkA - 1u) >> kChunkALog2;
            uint32_t cinc = c1n + c2n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, cinc, o);
                if (lane >= o) cinc += t;
            }
            s_cpre[2 * lane] = cinc - c1n - c2n;
            s_cpre[2 * lane + 1] = cinc - c2n;
            if (lane == 31) s_cpre[2 * kMaxRows] = cinc;
            s_mlo[lane] = rr.mlo;
            s_mcnt[lane] = rr.mcnt;
            if (lane == 0) { s_rows = rows; s_cy0 = tc.cy_lo; }
        }
        __syncthreads();
        const Cand32 *__restrict__ pts = sorted_q + sd.off;
        const uint32_t gtot = s_cpre[2 * kMaxRows];         // kChunkA-candidate chunks of the tile, dealt round-robin to the warps
        // this part's chunks of the flattened sequence (a tile that is not split is one part), and in them this
        // warp's share of every run, compacted into the warp's own run list (two run slots per lane: left / right
        // of the moment cells)
        const uint32_t pc0 = (uint32_t)((uint64_t)gtot * part_k / part_K), pc1 = (uint32_t)((uint64_t)gtot * (part_k + 1u) / part_K);
        uint32_t my_trips = 0u;
        int my_runs = 0;
        {
            uint32_t *ws = s_wstart[warp], *wl = s_wlen[warp];
            uint32_t st2[2], ln2[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int e = 2 * lane + k;
                part_run<kChunkALog2>(s_cpre[e], s_cpre[e + 1], s_rs[e], s_len[e], pc0, pc1, warp, st2[k], ln2[k]);
                my_trips += (ln2[k] + kStrideA - 1u) / kStrideA;
            }
            const unsigned m0 = __ballot_sync(0xffffffffu, ln2[0] > 0u), m1 = __ballot_sync(0xffffffffu, ln2[1] > 0u);
            const unsigned lt = (1u << lane) - 1u;
            const int k0 = __popc(m0 & lt) + __popc(m1 & lt);
            if (ln2[0] > 0u) { ws[k0] = st2[0]; wl[k0] = ln2[0]; }
            if (ln2[1] > 0u) { const int k1 = k0 + (ln2[0] > 0u ? 1 : 0); ws[k1] = st2[1]; wl[k1] = ln2[1]; }
            my_runs = __popc(m0) + __popc(m1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_trips += __shfl_xor_sync(0xffffffffu, my_trips, o);
            __syncwarp();
        }
        double acc64[kTile][kTile];
        walk_tile32<kStrideA, true, kPairA>(pts, s_wstart[warp], s_wlen[warp], my_runs, my_trips, lane, org,
                                            far_candidate(org, g), f1, f2, f3, 1.0f, acc64);
        uint32_t ncand = gtot * kChunkA;                      // upper bound of the candidates (last chunks are ragged)
        if (MOMENTS && part_k == 0u) {                        // the moment cells go with the first part
            const double *__restrict__ mom = moments + (size_t)set * ncells * kMomentDoubles;
            const int rows = s_rows, cy0 = s_cy0;
            for (int rr = 0; rr < rows; ++rr) {
                const int mc = s_mcnt[rr], cy = cy0 + rr;
                ncand += (uint32_t)mc;
                for (int k = threadIdx.x; k < mc; k += kDensThreads) {
                    const int cx = s_mlo[rr] + k;
                    accumulate_cell(mom + (size_t)(cy * g.C + cx) * kMomentDoubles, cell_centre(g, cx),
                                    cell_centre(g, cy), gx0, gy0, g.inv_h, c1, c2, c3, acc64);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) s_red[(i * kTile + j) * kRedStride + threadIdx.x] = acc64[i][j];
        __syncthreads();
        const int node = threadIdx.x >> 3, part = threadIdx.x & 7;
        double S = 0.0;
#pragma unroll
        for (int k = 0; k < kDensThreads / 8; ++k) S += s_red[node * kRedStride + part + 8 * k];
        S += __shfl_xor_sync(0xffffffffu, S, 1);
        S += __shfl_xor_sync(0xffffffffu, S, 2);
        S += __shfl_xor_sync(0xffffffffu, S, 4);
        if (stats && threadIdx.x == 0) { atomicAdd(&stats[0], part_k == 0u ? 1ull : 0ull); atomicAdd(&stats[2], (unsigned long long)(pc1 - pc0)); }
        if (!combine_parts(sb, part_k, part_K, pbase, part == 0, node, S, &s_ticket)) continue;   // not the last part
        const int ix = ix0 + node / kTile, iy = iy0 + node % kTile;
        const bool valid = ix < g.R && iy < g.R;
        if (MOMENTS && part_k != 0u)                              // thr counts the whole tile's candidates
            for (int rr = 0; rr < s_rows; ++rr) ncand += (uint32_t)s_mcnt[rr];
        const double thr = fmax((double)kBias32 * (double)max(ncand, 1024u), (double)min_thr);
        const int low = __syncthreads_or(valid && S < thr);      // also: everyone is done reading s_red
        if (low) {                                               // cannot certify a node here: fp64 does this tile
            if (threadIdx.x == 0) {
                if (stats) atomicAdd(&stats[4], 1ull);
                // the fp64 kernel walks it in the same parts (the part slots are free again: the ticket was reset)
                const unsigned slot = atomicAdd(&redo_counters[list], part_K);
                for (uint32_t q = 0; q < part_K; ++q) {
                    redo_list[(size_t)list * capacity + slot + q] = entry;
                    redo_partinfo[(size_t)(list / kNumClasses) * capacity + slot + q] = make_uint2(q | (part_K << 8), pbase);
                }
            }
            continue;
        }
        double P = 0.0;
        if (part == 0 && valid) {
            const double d = S * sd.knorm_over_n;
            P = gamma_eps(d, g);
            dens[(size_t)set * G + (size_t)ix * g.R + iy] = d;
            pz[(size_t)set * G + (size_t)ix * g.R + iy] = P;
        }
        P += __shfl_xor_sync(0xffffffffu, P, 8);
        P += __shfl_xor_sync(0xffffffffu, P, 16);
        if (lane == 0) s_red[warp] = P;
        __syncthreads();
        if (threadIdx.x == 0)
            tilesum[(size_t)set * g.NT * g.NT + tile] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
        __syncthreads();
    }

    // ---------------- phase B: warp per tile ----------------
    double *__restrict__ wred = s_red + warp * (kTile * kTile * kWarpRedStride);
    uint32_t *wstart = s_wstart[warp], *wlen = s_wlen[warp];
    __syncthreads();
    for (;;) {
        unsigned item = 0u;
        if (lane == 0) item = atomicAdd(&counters[kTicketB], 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nB) break;
        int list;
        unsigned index;
        const uint32_t entry = list_entry_ex(worklist, capacity, item, s_preB, s_listB,
                                             kMaxGroups * (kNumClasses - 1), list, index);
        const int set = (int)(entry >> 24), tile = (int)(entry & 0xffffffu);
        const double knorm_over_n = wd.set[set].knorm_over_n;
        const Cand32 *__restrict__ pts = sorted_q + wd.set[set].off;
        const int tx = tile / g.NT, ty = tile % g.NT;
        const int ix0 = tx * kTile, iy0 = ty * kTile;
        const double gx0 = (double)ix0 * g.step + g.gmin, gy0 = (double)iy0 * g.step + g.gmin;
        const TileOrigin32 org{kFp32FromDoubles ? 0u : quantise(gx0, g), kFp32FromDoubles ? 0u : quantise(gy0, g), gx0, gy0, g.inv_h, unit};
        const uint32_t *__restrict__ cs = cellstart + (size_t)set * (ncells + 1);
        const TileCells tc = tile_cells(g, ix0, iy0);
        const int rows = tc.cy_hi - tc.cy_lo + 1;
        RowRuns rr = {0u, 0u, 0u, 0u, 0, 0};
        if (lane < rows) rr = tile_row_runs(g, tc, cs, tc.cy_lo + lane, MOMENTS);
        // compact the non-empty runs, in (row, part) order, into this warp's shared arrays
        const unsigned m1 = __ballot_sync(0xffffffffu, rr.l1 > 0u);
        const unsigned m2 = MOMENTS ? __ballot_sync(0xffffffffu, rr.l2 > 0u) : 0u;
        const unsigned lt = (1u << lane) - 1u;
        __syncwarp();
        if (rr.l1 > 0u) { const int k = __popc(m1 & lt) + __popc(m2 & lt); wstart[k] = rr.s1; wlen[k] = rr.l1; }
        if (MOMENTS && rr.l2 > 0u) { const int k = __popc(m1 & lt) + __popc(m2 & lt) + (rr.l1 > 0u ? 1 : 0); wstart[k] = rr.s2; wlen[k] = rr.l2; }
        const int nruns = __popc(m1) + __popc(m2);
        uint32_t total_trips = ((rr.l1 + 63u) >> 6) + ((rr.l2 + 63u) >> 6);      // 64 candidates per trip here
        uint32_t ncand = rr.l1 + rr.l2 + (uint32_t)rr.mcnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            total_trips += __shfl_xor_sync(0xffffffffu, total_trips, o);
            ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
        }
        __syncwarp();
        const double thr = fmax((double)kBias32 * (double)max(ncand, 1024u), (double)min_thr);
        const int node = lane >> 1, half = lane & 1;
        const int ix = ix0 + node / kTile, iy = iy0 + node % kTile;
        const bool valid = ix < g.R && iy < g.R;
        float one = 1.0f;
        double S = 0.0, S1 = 0.0;
        bool requeue = false;
#pragma unroll 1
        for (int pass = 0;; ++pass) {
            double acc64[kTile][kTile];
            walk_tile32<64, false, true>(pts, wstart, wlen, nruns, total_trips, lane, org, far_candidate(org, g), f1, f2, f3, one, acc64);
            if (MOMENTS) {
                const double *__restrict__ mom = moments + (size_t)set * ncells * kMomentDoubles;
                for (int r = 0; r < rows; ++r) {
                    const int mc = __shfl_sync(0xffffffffu, rr.mcnt, r);
                    const int mlo = __shfl_sync(0xffffffffu, rr.mlo, r);
                    const int cy = tc.cy_lo + r;
                    for (int k = lane; k < mc; k += 32)
                        accumulate_cell(mom + (size_t)(cy * g.C + mlo + k) * kMomentDoubles, cell_centre(g, mlo + k),
                                        cell_centre(g, cy), gx0, gy0, g.inv_h, c1, c2, c3, acc64);
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < kTile; ++i)
#pragma unroll
                for (int j = 0; j < kTile; ++j) wred[(i * kTile + j) * kWarpRedStride + lane] = acc64[i][j];
            __syncwarp();
            S = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) S += wred[node * kWarpRedStride + half + 2 * k];
            S += __shfl_xor_sync(0xffffffffu, S, 1);
            if (stats && lane == 0) {
                atomicAdd(&stats[pass == 0 ? 0 : 1], 1ull);
                atomicAdd(&stats[pass == 0 ? 2 : 3], (unsigned long long)total_trips);
            }
            if (pass == 0) {
                if (!__any_sync(0xffffffffu, valid && S < thr)) break;      // every node certainly non-zero
                S1 = S;
                one = 1.0f + kBias32;                                       // second walk, every a_i raised by the bias
            } else {
                requeue = __any_sync(0xffffffffu, valid && S1 < thr && S > 0.0);
                S = (S1 < thr) ? 0.0 : S1;                                  // biased sum exactly 0: certainly no in-support point
                break;
            }
        }
        if (requeue) {
            if (lane == 0) {
                if (stats) atomicAdd(&stats[4], 1ull);
                const unsigned slot = atomicAdd(&redo_counters[list], 1u);
                redo_list[(size_t)list * capacity + slot] = entry;
            }
            continue;
        }
        double P = 0.0;
        if (half == 0 && valid) {
            const double d = S * knorm_over_n;
            P = gamma_eps(d, g);
            dens[(size_t)set * G + (size_t)ix * g.R + iy] = d;
            pz[(size_t)set * G + (size_t)ix * g.R + iy] = P;
        }
#pragma unroll
        for (int o = 2; o < 32; o <<= 1) P += __shfl_xor_sync(0xffffffffu, P, o);
        if (lane == 0) tilesum[(size_t)set * g.NT * g.NT + tile] = P;
    }
}

// ------------------------------------------------------------------------------------------
// k_density_dense64: brute force in fp64, every node against every point (cross-check path and the
// gaussian kernel).  grid (tiles of 32x32 nodes, n_sets, point slices); 256 threads, 2x2 nodes per
// thread; points staged through shared memory.  Slice partials go to `partial[slice][set][G]`.
constexpr int kDenseTile = 32;
constexpr int kDenseChunk = 512;
template <int KERNEL>
__global__ void __launch_bounds__(256)
k_density_dense64(const __grid_constant__ WaveDesc wd, const GridGeom g,
                  const double2 *__restrict__ pts_all, double *__restrict__ partial, int nslices) {
    __shared__ double2 s_p[kDenseChunk];
    const int set = blockIdx.y;
    const SetDesc sd = wd.set[set];
    const int ntd = (g.R + kDenseTile - 1) / kDenseTile;
    const int tx = blockIdx.x / ntd, ty = blockIdx.x % ntd;
    const int ix = tx * kDenseTile + (threadIdx.x / 16) * 2, iy = ty * kDenseTile + (threadIdx.x % 16) * 2;
    const double gx[2] = {(double)ix * g.step + g.gmin, (double)(ix + 1) * g.step + g.gmin};
    const double gy[2] = {(double)iy * g.step + g.gmin, (double)(iy + 1) * g.step + g.gmin};
    const double2 *__restrict__ pts = pts_all + sd.off;
    const int per = (sd.n + nslices - 1) / nslices;
    const int beg = blockIdx.z * per, end = min(beg + per, sd.n);
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int c0 = beg; c0 < end; c0 += kDenseChunk) {
        const int m = min(kDenseChunk, end - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += 256) s_p[i] = pts[c0 + i];
        __syncthreads();
        for (int i = 0; i < m; ++i) {
            const double2 p = s_p[i];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const double ux = (gx[a] - p.x) * g.inv_h, uy = (gy[b] - p.y) * g.inv_h;
                    if (KERNEL == 0) {
                        const double t = 1.0 - (ux * ux + uy * uy);
                        if (t > 0.0) acc[a][b] += t;
                    } else {
                        acc[a][b] += exp(-0.5 * (ux * ux + uy * uy));
                    }
                }
        }
    }
    const size_t G = (size_t)g.R * g.R;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
            if (ix + a < g.R && iy + b < g.R)
                partial[((size_t)blockIdx.z * wd.n_sets + set) * G + (size_t)(ix + a) * g.R + iy + b] = acc[a][b];
}

// ------------------------------------------------------------------------------------------
// k_density_dense32: the brute-force N x G kernel of the north star -- every node against every point,
// fp32 terms on the FP32 (FMA) pipe.  grid (node blocks of 128x64, n_sets, point slices); 128 threads;
// each thread keeps an 8x8 patch of nodes in 64 fp32 registers.  A slice (<= kD32Slice points) is staged
// once in shared memory as fp64 and broadcast to all threads; every thread re-centres a point on ITS OWN
// first node in fp64 (2 DADD + 2 DMUL on the otherwise idle fp64 pipe) before rounding to fp32, so the
// fp32 coordinates are small (|u| < ~2 for any pair that can be in support) and the per-term error stays
// ~1e-7.  Per pair: FFMA.SAT (t = sat(a_i - v_j^2): an exact ReLU because t <= 1) + FADD, i.e. 2
// FMA-pipe lane-ops; 24 more per point per thread -> 152 per 64 pairs.  fp32 partial sums never run
// longer than one slice; slices are folded in fp64 by k_dense_finish.
// KERNEL 1 (gaussian): the term is separable, exp(-d^2/2) = ex_i * ey_j, so a point costs a thread 16
// MUFU.EX2 and each pair ONE FFMA (acc += ex_i * ey_j) -- FMA-pipe bound like the Epanechnikov variant, not
// SFU bound (an ex2 per pair would cap the kernel at 16 pairs/clk/SM, an eighth of the FFMA rate).
constexpr int kD32Threads = 128;
constexpr int kD32Patch = 8;
constexpr int kD32TilesX = 16, kD32TilesY = 8;                 // threads per block along x / y
constexpr int kD32NodesX = kD32TilesX * kD32Patch;             // 128
constexpr int kD32NodesY = kD32TilesY * kD32Patch;             // 64
constexpr int kD32Slice = 1024;
template <int KERNEL>
__global__ void __launch_bounds__(kD32Threads)
k_density_dense32(const __grid_constant__ WaveDesc wd, const GridGeom g,
                  const double2 *__restrict__ pts_all, double *__restrict__ partial) {
    __shared__ double2 s_p[kD32Slice];
    const int set = blockIdx.y;
    const SetDesc sd = wd.set[set];
    const int beg = blockIdx.z * kD32Slice;
    const size_t G = (size_t)g.R * g.R;
    const int nby = (g.R + kD32NodesY - 1) / kD32NodesY;
    const int bx = blockIdx.x / nby, by = blockIdx.x % nby;
    const int ix0 = bx * kD32NodesX + (threadIdx.x / kD32TilesY) * kD32Patch;
    const int iy0 = by * kD32NodesY + (threadIdx.x % kD32TilesY) * kD32Patch;
    float acc[kD32Patch][kD32Patch];
#pragma unroll
    for (int i = 0; i < kD32Patch; ++i)
#pragma unroll
        for (int j = 0; j < kD32Patch; ++j) acc[i][j] = 0.f;
    if (beg < sd.n) {
        const int m = min(kD32Slice, sd.n - beg);
        const double2 *__restrict__ pts = pts_all + sd.off + beg;
        for (int i = threadIdx.x; i < m; i += kD32Threads) s_p[i] = pts[i];
        __syncthreads();
        const double x0 = (double)ix0 * g.step + g.gmin, y0 = (double)iy0 * g.step + g.gmin;
        const float dh = (float)g.dh;
        const float kexp = 0.72134752044448170368f;            // 0.5 * log2(e)
#pragma unroll 2
        for (int k = 0; k < m; ++k) {
            const double2 p = s_p[k];
            const float u0 = (float)((p.x - x0) * g.inv_h);    // node i sits at i*dh - u0 (units of h)
            const float w0 = (float)((p.y - y0) * g.inv_h);
            float a[kD32Patch], v[kD32Patch];
#pragma unroll
            for (int i = 0; i < kD32Patch; ++i) {
                const float ux = (float)i * dh - u0;
                const float vy = (float)i * dh - w0;
                a[i] = (KERNEL == 0) ? fmaf(-ux, ux, 1.0f) : exp2f(-kexp * ux * ux);
                v[i] = (KERNEL == 0) ? vy : exp2f(-kexp * vy * vy);
            }
#pragma unroll
            for (int i = 0; i < kD32Patch; ++i)
#pragma unroll
                for (int j = 0; j < kD32Patch; ++j) {
                    if (KERNEL == 0) acc[i][j] += __saturatef(fmaf(-v[j], v[j], a[i]));
                    else acc[i][j] = fmaf(a[i], v[j], acc[i][j]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < kD32Patch; ++i)
#pragma unroll
        for (int j = 0; j < kD32Patch; ++j)
            if (ix0 + i < g.R && iy0 + j < g.R)
                partial[((size_t)blockIdx.z * wd.n_sets + set) * G + (size_t)(ix0 + i) * g.R + iy0 + j] = (double)acc[i][j];
}

// Fold the dense slice partials in slice order and apply the per-node epilogue; also produces the
// per-"tile" normaliser partials in the layout k_js expects (here one entry per 256 nodes).
__global__ void __launch_bounds__(256)
k_dense_finish(const __grid_constant__ WaveDesc wd, const GridGeom g, const double *__restrict__ partial,
               int nslices, double *__restrict__ dens, double *__restrict__ pz,
               double *__restrict__ tilesum, int nparts) {
    __shared__ double s_w[32];
    const int set = blockIdx.y;
    const size_t G = (size_t)g.R * g.R;
    const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
    double P = 0.0;
    if (k < G) {
        double S = 0.0;
        for (int s = 0; s < nslices; ++s) S += partial[((size_t)s * wd.n_sets + set) * G + k];
        const double d = S * wd.set[set].knorm_over_n;
        P = gamma_eps(d, g);
        dens[(size_t)set * G + k] = d;
        pz[(size_t)set * G + k] = P;
    }
    const double tot = block_sum(P, s_w);
    if (threadIdx.x == 0) tilesum[(size_t)set * nparts + blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------
// log(x/m).  Where x and m almost coincide it is taken as log1p((x-m)/m), which keeps the d^2/m term that
// log(x/m) loses to rounding (x - m is exact there); elsewhere it is the plain log of the quotient (log1p
// would lose x entirely once x/m < 1e-16, which gamma > 1 produces routinely).
__device__ __forceinline__ double log_ratio(double x, double m) {
    const double d = x - m;
    return (fabs(d) < 0.25 * m) ? log1p(d / m) : log(x / m);
}

// k_set_totals: grid (n_sets), 256 threads: a set's normaliser = its per-tile partials summed in a
// fixed order (KDE_distance.py:76, z.sum()).
__global__ void __launch_bounds__(kJsThreads)
k_set_totals(int nparts, const double *__restrict__ tilesum, double *__restrict__ totals) {
    __shared__ double s_w[32];
    const double *__restrict__ ts = tilesum + (size_t)blockIdx.x * nparts;
    double a = 0.0;
    for (int k = threadIdx.x; k < nparts; k += kJsThreads) a += ts[k];
    a = block_sum(a, s_w);
    if (threadIdx.x == 0) totals[blockIdx.x] = a;
}

// k_js: grid (node blocks, n_items), 256 threads.  Normalises with totals[set] (k_set_totals, or the
// all-reduced normalisers in row-sharded mode), accumulates its nodes' rel_entr(p,m) and rel_entr(q,m);
// the last block per item (ticket) folds the block partials in block order:
// js = sqrt((left + right) / 2)          scipy:spatial/distance.py:1278-1290.
__global__ void __launch_bounds__(kJsThreads)
k_js(const __grid_constant__ WaveDesc wd, int G, int nparts, int k_begin, int k_end,
     const double *__restrict__ pz, const double *__restrict__ totals,
     const RawStat *__restrict__ rstat,
     double *__restrict__ z_out, double *__restrict__ jspart, unsigned *__restrict__ ticket,
     double *__restrict__ out_js, int32_t *__restrict__ out_status, double *__restrict__ out_lr,
     double *__restrict__ out_norm, int item_base) {
    __shared__ double s_w[32];
    __shared__ bool s_last;
    const int item = blockIdx.y;
    const int s0 = 2 * item, s1 = 2 * item + 1;
    const double tp = totals[s0], tq = totals[s1];
    double left = 0.0, right = 0.0;
    const int kb = k_begin + blockIdx.x * kJsNodesPerBlock;
    const int ke = min(kb + kJsNodesPerBlock, k_end);
    const bool normal = (tp > 0.0) && (tq > 0.0);          // false for a zero or NaN normaliser (empty KDE)
    for (int k = kb + threadIdx.x; k < ke; k += kJsThreads) {
        const double a = pz[(size_t)s0 * G + k], b = pz[(size_t)s1 * G + k];
        if (normal && a == 0.0 && b == 0.0) {              // both probabilities are exactly 0: nothing to add
            if (z_out) { z_out[(size_t)s0 * G + k] = 0.0; z_out[(size_t)s1 * G + k] = 0.0; }
            continue;
        }
        const double p = a / tp;                           // z /= z.sum()  KDE_distance.py:76
        const double q = b / tq;
        const double m = (p + q) / 2.0;
        // rel_entr(p,m) + rel_entr(q,m) of ONE node is >= 0 (log-sum inequality) while the two separate sums
        // scipy forms cancel against each other: for near-identical distributions left ~ -right ~ 1e-3 and
        // left + right ~ 1e-17, whose sign is rounding noise (sqrt of it can come out NaN).  The per-node sum
        // goes into `left`; `right` stays 0, so left + right is the well-conditioned total.
        double node = 0.0;
        if (p > 0.0) node = p * log_ratio(p, m);           // rel_entr(0, m) = 0
        if (q > 0.0) node += q * log_ratio(q, m);
        left += node;
        if (p != p) left = p;                              // NaN normaliser (empty KDE) propagates
        if (q != q) left = q;
        if (z_out) { z_out[(size_t)s0 * G + k] = p; z_out[(size_t)s1 * G + k] = q; }
    }
    left = block_sum(left, s_w);                           // `right` is 0 in every thread
    if (threadIdx.x == 0) {
        jspart[((size_t)item * gridDim.x + blockIdx.x) * 2 + 0] = left;
        jspart[((size_t)item * gridDim.x + blockIdx.x) * 2 + 1] = right;
        __threadfence();
        unsigned t = atomicAdd(&ticket[item], 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {          // warp 0 of the last block folds the block partials, fixed shape
        __threadfence();
        const volatile double *vp = jspart + (size_t)item * gridDim.x * 2;
        double L = 0.0, Rr = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) { L += vp[2 * i]; Rr += vp[2 * i + 1]; }
        L = warp_sum(L);
        Rr = warp_sum(Rr);
        if (threadIdx.x == 0) {
            const int bad = (rstat[wd.set[s0].rs_self].flags | rstat[wd.set[s0].rs_other].flags) & 1;
            if (out_lr) { out_lr[0] = L; out_lr[1] = Rr; }
            if (out_norm) { out_norm[0] = tp; out_norm[1] = tq; }
            double js2 = (L + Rr) / 2.0;
            if (js2 < 0.0) js2 = 0.0;          // rounding can leave -1e-20 where the exact value is +1e-20; NaN stays NaN
            if (out_js) out_js[item_base + item] = bad ? __longlong_as_double(0x7ff8000000000000LL) : sqrt(js2);
            if (out_status) out_status[item_base + item] = bad ? 3 : 0;
            ticket[item] = 0;
        }
    }
}

// k_js_tiles: the divergence of the binned path when nobody asked for the grids.  grid (NT tile rows, n_items),
// 256 threads; a half-warp owns one 4x4-node tile at a time (its 16 nodes are four 32-byte segments of pz).
// Tiles without candidates were only FLAGGED by k_tile_classify (tileflag = 1), their node values never written:
// such a node's value is the constant gamma_eps(0), so neither kernel touches those ~70 % of the grid in memory.
// Arithmetic per node and the last-block fold are k_js's.
__global__ void __launch_bounds__(kJsThreads)
k_js_tiles(const __grid_constant__ WaveDesc wd, const GridGeom g, const double *__restrict__ pz,
           const double *__restrict__ totals, const RawStat *__restrict__ rstat,
           const uint8_t *__restrict__ tileflag, double *__restrict__ jspart, unsigned *__restrict__ ticket,
           double *__restrict__ out_js, int32_t *__restrict__ out_status, int item_base) {
    __shared__ double s_w[32];
    __shared__ bool s_last;
    const int item = blockIdx.y;
    const int s0 = 2 * item, s1 = 2 * item + 1;
    const double tp = totals[s0], tq = totals[s1];
    const bool normal = (tp > 0.0) && (tq > 0.0);
    const double P0 = gamma_eps(0.0, g);
    const size_t G = (size_t)g.R * g.R;
    const int ntiles = g.NT * g.NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = blockIdx.x, ix = tx * kTile + ((lane & 15) >> 2), jj = lane & 3;
    double left = 0.0;
    if (ix < g.R) {
        for (int ty = warp * 2 + (lane >> 4); ty < g.NT; ty += 2 * (kJsThreads / 32)) {
            const int iy = ty * kTile + jj;
            if (iy >= g.R) continue;
            const int tile = tx * g.NT + ty;
            const size_t k = (size_t)ix * g.R + iy;
            const double a = tileflag[(size_t)s0 * ntiles + tile] ? P0 : pz[(size_t)s0 * G + k];
            const double b = tileflag[(size_t)s1 * ntiles + tile] ? P0 : pz[(size_t)s1 * G + k];
            if (normal && a == 0.0 && b == 0.0) continue;
            const double p = a / tp;                           // z /= z.sum()  KDE_distance.py:76
            const double q = b / tq;
            const double m = (p + q) / 2.0;
            double node = 0.0;
            if (p > 0.0) node = p * log_ratio(p, m);           // rel_entr(0, m) = 0
            if (q > 0.0) node += q * log_ratio(q, m);
            left += node;
            if (p != p) left = p;                              // NaN normaliser (empty KDE) propagates
            if (q != q) left = q;
        }
    }
    left = block_sum(left, s_w);
    if (threadIdx.x == 0) {
        jspart[((size_t)item * gridDim.x + blockIdx.x) * 2 + 0] = left;
        jspart[((size_t)item * gridDim.x + blockIdx.x) * 2 + 1] = 0.0;
        __threadfence();
        unsigned t = atomicAdd(&ticket[item], 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {          // warp 0 of the last block folds the block partials, fixed shape
        __threadfence();
        const volatile double *vp = jspart + (size_t)item * gridDim.x * 2;
        double L = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) L += vp[2 * i];
        L = warp_sum(L);
        if (threadIdx.x == 0) {
            const int bad = (rstat[wd.set[s0].rs_self].flags | rstat[wd.set[s0].rs_other].flags) & 1;
            double js2 = L / 2.0;
            if (js2 < 0.0) js2 = 0.0;          // rounding can leave -1e-20 where the exact value is +1e-20; NaN stays NaN
            if (out_js) out_js[item_base + item] = bad ? __longlong_as_double(0x7ff8000000000000LL) : sqrt(js2);
            if (out_status) out_status[item_base + item] = bad ? 3 : 0;
            ticket[item] = 0;
        }
    }
}

// Sum of `count` per-tile normaliser partials of both sets of one item (row-sharded phase 1):
// 1 block; set 1's partials start `stride` doubles after set 0's.
__global__ void __launch_bounds__(kJsThreads)
k_norm_totals(int count, int stride, const double *__restrict__ tilesum, double *__restrict__ out_norm) {
    __shared__ double s_w[32];
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < count; k += kJsThreads) { a += tilesum[k]; b += tilesum[stride + k]; }
    a = block_sum(a, s_w);
    b = block_sum(b, s_w);
    if (threadIdx.x == 0) { out_norm[0] = a; out_norm[1] = b; }
}


// ==========================================================================================
// Band mode: ONE huge evaluation over several GPUs (config C5).  The grid is cut along column 1 (tile columns ty)
// into one band per rank; cell rows of the sort run along the same axis, so "every point within h of a band" is a
// contiguous slice of a rank's locally sorted points.  Per rank:
//   k_minmax (own slices) -> all-gather of the extrema -> k_band_fold_stats (joint scaler, identical on all ranks)
//   local sort of the own slices (k_scale_count / k_scan / k_place) -> all-gather of the per-cell counts
//   k_band_cells (global counts: work per tile column -> band cuts; local counts of the rows a band needs)
//   point exchange (NCCL all-to-all of contiguous slices) -> k_band_merge (per cell: sources in rank order ==
//   original index order, so the band's sorted array is exactly the single-GPU one restricted to its rows)
//   moments / classify / density on the band's tiles -> k_band_norm_totals -> all-gather (2 doubles)
//   k_band_fold_norms -> k_js_tiles over the band -> all-gather (2 doubles) -> sqrt on the host.

// Joint extrema of the two sets from every rank's partial extrema (rank order; min/max are order-independent).
__global__ void k_band_fold_stats(const RawStat *__restrict__ all, int world, RawStat *__restrict__ out) {
    const int set = threadIdx.x;
    if (set >= 2) return;
    RawStat a = all[set];
    for (int r = 1; r < world; ++r) {
        const RawStat b = all[(size_t)r * 2 + set];
        a.mn[0] = fmin(a.mn[0], b.mn[0]); a.mx[0] = fmax(a.mx[0], b.mx[0]);
        a.mn[1] = fmin(a.mn[1], b.mn[1]); a.mx[1] = fmax(a.mx[1], b.mx[1]);
        a.flags |= b.flags;
    }
    a.pad = 0;
    out[set] = a;
}

// Row starts of a locally sorted set: rowstart[set][cy] = cellstart[set][cy * C], cy = 0..C (appended to the
// per-rank count block that is all-gathered, so the host can size the exchange from a few KB).
__global__ void k_band_rowstarts(const uint32_t *__restrict__ cellstart, int C, uint32_t *__restrict__ rowstart) {
    const int set = blockIdx.x, ncells = C * C;
    for (int cy = threadIdx.x; cy <= C; cy += blockDim.x)
        rowstart[(size_t)set * (C + 1) + cy] = cellstart[(size_t)set * (ncells + 1) + (size_t)cy * C];
}

// k_band_cell_totals: grid (ceil(ncells / 256), 2), 256 threads.  totals[set][c] = sum over ranks of count_r(set, c)
// for cells in [c_lo, c_hi), 0 for all others.  cs_all: per rank a block of `rank_stride` u32 starting with
// cellstart_r[2][ncells + 1].
__global__ void __launch_bounds__(256)
k_band_cell_totals(const uint32_t *__restrict__ cs_all, int world, size_t rank_stride, int ncells, int c_lo, int c_hi,
                   uint32_t *__restrict__ totals) {
    const int set = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
    if (c >= ncells) return;
    uint32_t t = 0u;
    if (c >= c_lo && c < c_hi)
        for (int r = 0; r < world; ++r) {
            const uint32_t *cs = cs_all + (size_t)r * rank_stride + (size_t)set * (ncells + 1);
            t += cs[c + 1] - cs[c];
        }
    totals[(size_t)set * ncells + c] = t;
}

// k_band_cells: grid (2), 1024 threads.  cellstart_out[set][0..ncells] = exclusive scan of totals[set][].  With
// [0, ncells) in k_band_cell_totals this is the GLOBAL cell table (the same numbers the single-GPU sort produces),
// with a band's rows it is the band's local table.
__global__ void __launch_bounds__(1024)
k_band_cells(const uint32_t *__restrict__ totals, int ncells, uint32_t *__restrict__ cellstart_out) {
    const int set = blockIdx.x;
    constexpr int kCptMax = kMaxCells / 1024;
    const int cpt = (ncells + 1023) / 1024;
    const int c0 = threadIdx.x * cpt;
    uint32_t tot[kCptMax];
    uint32_t tsum = 0u;
#pragma unroll
    for (int k = 0; k < kCptMax; ++k) {
        const int c = c0 + k;
        tot[k] = (k < cpt && c < ncells) ? totals[(size_t)set * ncells + c] : 0u;
        tsum += tot[k];
    }
    __shared__ uint32_t s_w[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t v = s_w[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_w[lane] = iv - v;
    }
    __syncthreads();
    uint32_t run = s_w[w] + inc - tsum;
    uint32_t *out = cellstart_out + (size_t)set * (ncells + 1);
#pragma unroll
    for (int k = 0; k < kCptMax; ++k) {
        if (k < cpt && c0 + k < ncells) out[c0 + k] = run;
        run += tot[k];
    }
    if (threadIdx.x == 1023) out[ncells] = run;
}

// k_band_merge: grid (ceil(points of the larger set / (256 * kMergePer)), 2 sets), 256 threads: one thread per point
// of the band's sorted array.  Destination slot i of a set belongs to the cell c with cellstart_local[c] <= i <
// cellstart_local[c + 1] (binary search; neighbouring threads share the answer) and, inside the cell, to the first
// source rank whose piece reaches that far -- rank order is original index order, so the result equals the
// single-GPU stable sort restricted to these rows.  Real distance clouds are concentrated (a few cells hold most of
// the points): dealing out destination slots, not cells, keeps every thread busy, and kMergePer independent loads
// per thread keep enough bytes in flight for NVLink latency.  Also writes the fixed-point copy for the fp32 kernel.
// Where a source's points come from:
//   PEER = false  a staging buffer per set that holds, source after source, that rank's sorted points of cell rows
//                 [cy_lo, cy_hi] (filled by the NCCL point-to-point exchange)
//   PEER = true   the source rank's OWN locally sorted array (set A, then set B), mapped into this process over
//                 NVLink (kdejs_band_peer_*): the copy IS the exchange -- no staging buffer, no second pass
constexpr int kMergePer = 4;
constexpr int kMaxPeers = 16;
struct PeerPtrs { const double2 *p[kMaxPeers]; };
template <bool PEER>
__global__ void __launch_bounds__(256)
k_band_merge(const double2 *__restrict__ stage_a, const double2 *__restrict__ stage_b, const PeerPtrs peers,
             const uint32_t *__restrict__ cs_all, int world, size_t rank_stride, const GridGeom g, int cy_lo, int cy_hi,
             const uint32_t *__restrict__ cellstart_local, int64_t off_b,
             double2 *__restrict__ sorted, uint2 *__restrict__ sorted_q) {
    __shared__ uint32_t s_srcbase[kMaxPeers];     // PEER: first point of the set in rank r's array; else in the staging buffer
    const int set = blockIdx.y;
    const int ncells = g.C * g.C;
    const int c_lo = cy_lo * g.C, c_hi = (cy_hi + 1) * g.C;
    const uint32_t *__restrict__ csl = cellstart_local + (size_t)set * (ncells + 1);
    const uint32_t n_set = csl[c_hi];
    if ((uint32_t)blockIdx.x * (256u * kMergePer) >= n_set) return;
    if (threadIdx.x == 0) {
        uint32_t run = 0u;
        for (int r = 0; r < world; ++r) {
            const uint32_t *cs = cs_all + (size_t)r * rank_stride + (size_t)set * (ncells + 1);
            if (PEER) s_srcbase[r] = set ? cs_all[(size_t)r * rank_stride + ncells] : 0u;
            else { s_srcbase[r] = run - cs[c_lo]; run += cs[c_hi] - cs[c_lo]; }      // (mod 2^32: cs[cell] is added back below)
        }
    }
    __syncthreads();
    const double2 *__restrict__ stage = set ? stage_b : stage_a;
    const int64_t base = set ? off_b : 0;
    const double2 *src[kMergePer];
#pragma unroll
    for (int u = 0; u < kMergePer; ++u) {
        const uint32_t i = ((uint32_t)blockIdx.x * kMergePer + u) * 256u + threadIdx.x;
        src[u] = nullptr;
        if (i >= n_set) continue;
        int lo = c_lo, hi = c_hi;                 // csl[lo] <= i < csl[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (csl[mid] <= i) lo = mid; else hi = mid;
        }
        uint32_t o = i - csl[lo];
        int r = 0;
        uint32_t b = 0u;
        for (; r < world; ++r) {
            const uint32_t *cs = cs_all + (size_t)r * rank_stride + (size_t)set * (ncells + 1);
            b = cs[lo];
            const uint32_t cnt = cs[lo + 1] - b;
            if (o < cnt) break;
            o -= cnt;
        }
        src[u] = (PEER ? peers.p[r] : stage) + (uint32_t)(s_srcbase[r] + b + o);
    }
    double2 p[kMergePer];
#pragma unroll
    for (int u = 0; u < kMergePer; ++u)          // a peer rewrites its array between evaluations: no stale lines
        if (src[u]) p[u] = PEER ? __ldcv(src[u]) : *src[u];
#pragma unroll
    for (int u = 0; u < kMergePer; ++u) {
        const uint32_t i = ((uint32_t)blockIdx.x * kMergePer + u) * 256u + threadIdx.x;
        if (!src[u]) continue;
        sorted[base + i] = p[u];
        if (sorted_q) sorted_q[base + i] = make_uint2(quantise(p[u].x, g), quantise(p[u].y, g));
    }
}

// Estimated density work of one tile COLUMN (all tile rows, both sets): the band cuts are derived from these.
// moments_min_points >= 0: the moment-cell rule of the single-GPU path is applied here to the WHOLE evaluation
// (eligible geometry and at least that many points in total, read from the global cell table), and the decision is
// left in colwork[NT] so that the host learns it with the same copy.
__global__ void __launch_bounds__(512)
k_tile_colwork(const __grid_constant__ WaveDesc wd, const GridGeom g_in,
               const uint32_t *__restrict__ cellstart, double *__restrict__ colwork, long long moments_min_points) {
    __shared__ double s_w[32];
    GridGeom g = g_in;
    const int ncells = g.C * g.C;
    if (moments_min_points >= 0 && g.use_moments == 1) {
        long long total = 0;
        for (int s = 0; s < wd.n_sets; ++s) total += cellstart[(size_t)s * (ncells + 1) + ncells];
        if (total < moments_min_points) g.use_moments = 0;
    }
    const int ty = blockIdx.x;
    double work = 0.0;
    for (int idx = threadIdx.x; idx < wd.n_sets * g.NT; idx += blockDim.x) {
        const int set = idx / g.NT, tx = idx % g.NT;
        const TileCells tc = tile_cells(g, tx * kTile, ty * kTile);
        const uint32_t *__restrict__ cs = cellstart + (size_t)set * (ncells + 1);
        uint32_t total = 0u;
        for (int cy = tc.cy_lo; cy <= tc.cy_hi; ++cy) {
            const RowRuns rr = tile_row_runs(g, tc, cs, cy, g.use_moments != 0);
            total += rr.l1 + rr.l2 + (uint32_t)rr.mcnt;
        }
        work += (double)total + (total ? 64.0 : 2.0);
    }
    work = block_sum(work, s_w);
    if (threadIdx.x == 0) {
        colwork[ty] = work;
        if (ty == 0) colwork[g.NT] = g.use_moments ? 1.0 : 0.0;
    }
}

// Partial normalisers of a band: per-tile partials of tile columns [ty_b, ty_e), all tile rows, summed in a fixed
// order (thread-strided over the band's tiles in (tx, ty) order, then the block tree).  1 block.
__global__ void __launch_bounds__(kJsThreads)
k_band_norm_totals(const double *__restrict__ tilesum, int NT, int ty_b, int ty_e, double *__restrict__ out_norm) {
    __shared__ double s_w[32];
    const int nty = ty_e - ty_b, count = NT * nty;
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < count; k += kJsThreads) {
        const int tile = (k / nty) * NT + ty_b + k % nty;
        a += tilesum[tile];
        b += tilesum[(size_t)NT * NT + tile];
    }
    a = block_sum(a, s_w);
    b = block_sum(b, s_w);
    if (threadIdx.x == 0) { out_norm[0] = a; out_norm[1] = b; }
}

// Normalisers of the whole grid = the ranks' partials added in rank order (every rank gets the same bits).
__global__ void k_band_fold_norms(const double *__restrict__ norms_all, int world, double *__restrict__ totals) {
    if (threadIdx.x >= 2) return;
    double t = 0.0;
    for (int r = 0; r < world; ++r) t += norms_all[2 * r + threadIdx.x];
    totals[threadIdx.x] = t;
}

// Divergence partial of a band: k_js_tiles' arithmetic over tile columns [ty_b, ty_e) of one item, every node
// value read from pz (band mode writes empty tiles too).  grid (NT tile rows), 256 threads; the last block folds
// the block partials in block order into out_lr[0] ([1] = 0, [2] = 1 if an input held an infinity).
__global__ void __launch_bounds__(kJsThreads)
k_band_js(const GridGeom g, int ty_b, int ty_e, const double *__restrict__ pz, const double *__restrict__ totals,
          const RawStat *__restrict__ rstat, double *__restrict__ jspart, unsigned *__restrict__ ticket,
          double *__restrict__ out_lr) {
    __shared__ double s_w[32];
    __shared__ bool s_last;
    const double tp = totals[0], tq = totals[1];
    const bool normal = (tp > 0.0) && (tq > 0.0);
    const size_t G = (size_t)g.R * g.R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = blockIdx.x, ix = tx * kTile + ((lane & 15) >> 2), jj = lane & 3;
    double left = 0.0;
    if (ix < g.R) {
        for (int ty = ty_b + warp * 2 + (lane >> 4); ty < ty_e; ty += 2 * (kJsThreads / 32)) {
            const int iy = ty * kTile + jj;
            if (iy >= g.R) continue;
            const size_t k = (size_t)ix * g.R + iy;
            const double a = pz[k], b = pz[G + k];
            if (normal && a == 0.0 && b == 0.0) continue;
            const double p = a / tp, q = b / tq;
            const double m = (p + q) / 2.0;
            double node = 0.0;
            if (p > 0.0) node = p * log_ratio(p, m);
            if (q > 0.0) node += q * log_ratio(q, m);
            left += node;
            if (p != p) left = p;
            if (q != q) left = q;
        }
    }
    left = block_sum(left, s_w);
    if (threadIdx.x == 0) {
        jspart[blockIdx.x] = left;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        const volatile double *vp = jspart;
        double L = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) L += vp[i];
        L = warp_sum(L);
        if (threadIdx.x == 0) {
            out_lr[0] = L;
            out_lr[1] = 0.0;
            out_lr[2] = (double)((rstat[0].flags | rstat[1].flags) & 1);
            *ticket = 0u;
        }
    }
}


// ------------------------------------------------------------------------------------------
// k_summary: process_result's summary vector and the two ELFI nodes built on it, fused behind the JS reduction
// (poppunk_mod.py:285-291 calculate_gene_freqs, :379 [divergence, freq_core, freq_intermediate, freq_rare],
// :571-572 Distance("euclidean") against the observed vector and log).  grid (B), 256 threads; block b counts its
// simulation's gene frequencies X in exact integers and thread 0 writes out[b] = {JS, fc, fi, fr, d, log d}.
// Divisions are int / int in fp64 exactly as numpy's; the squared differences are summed in index order with
// separately rounded multiplies and adds (scipy's euclidean loop; no FMA contraction).
__global__ void __launch_bounds__(256)
k_summary(const double *__restrict__ freqs, const int64_t *__restrict__ offs, double core_thr, double rare_thr,
          const double *__restrict__ js, double o0, double o1, double o2, double o3, double *__restrict__ out) {
    __shared__ unsigned long long s_cnt[4];
    const int b = blockIdx.x;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned n_pos = 0u, n_core = 0u, n_rare = 0u, n_mid = 0u;
    for (int64_t k = offs[b] + threadIdx.x; k < offs[b + 1]; k += blockDim.x) {
        const double X = freqs[k];
        n_pos += (X > 0.0);
        n_core += (X >= core_thr);
        n_rare += (0.0 < X) && (X < rare_thr);
        n_mid += (rare_thr <= X) && (X < core_thr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
        n_core += __shfl_xor_sync(0xffffffffu, n_core, o);
        n_rare += __shfl_xor_sync(0xffffffffu, n_rare, o);
        n_mid += __shfl_xor_sync(0xffffffffu, n_mid, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[0], (unsigned long long)n_pos);
        atomicAdd(&s_cnt[1], (unsigned long long)n_core);
        atomicAdd(&s_cnt[2], (unsigned long long)n_rare);
        atomicAdd(&s_cnt[3], (unsigned long long)n_mid);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double genes = (double)s_cnt[0];
        const double fc = (double)s_cnt[1] / genes, fr = (double)s_cnt[2] / genes, fi = (double)s_cnt[3] / genes;
        const double v[4] = {js[b], fc, fi, fr}, o[4] = {o0, o1, o2, o3};
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double df = __dsub_rn(v[k], o[k]);
            acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
        const double d = sqrt(acc);
        double *r = out + (size_t)b * 6;
        r[0] = v[0]; r[1] = fc; r[2] = fi; r[3] = fr; r[4] = d; r[5] = log(d);
    }
}

// ------------------------------------------------------------------------------------------
// Issue-rate microbenchmarks (roofline denominators measured on the box, not quoted).
template <int WHICH>
__global__ void __launch_bounds__(256) k_peak(int iters, float *sink_f, double *sink_d) {
    if (WHICH == 0) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 1.0f + threadIdx.x * 1e-6f + i;
        const float b = 0.999999f, c = 1e-7f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678f) sink_f[0] = s;
    } else if (WHICH == 1) {
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i;
        const double b = 0.999999, c = 1e-7;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678) sink_d[0] = s;
    } else if (WHICH == 3 || WHICH == 4) {
        // DFMA stream with WHICH==3: one, WHICH==4: two independent FFMA (full-rate FP32 pipe) per DFMA.
        // Measured on B200: 36.5 TFLOP/s alone, 34.7 with one FFMA per DFMA, 22.4 with two -- i.e. the SMSP
        // issues one instruction per cycle and a second instruction fits beside each 2-cycle fp64 issue for
        // free; an fp64-bound loop therefore tolerates about one non-fp64 instruction per fp64 instruction.
        double a[8];
        float f[8], h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = 1.0 + threadIdx.x * 1e-6 + i; f[i] = 1.0f + i; h[i] = 2.0f + i; }
        const double b = 0.999999, c = 1e-7;
        const float bf = 0.999999f, cf = 1e-7f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fma(a[i], b, c);
                f[i] = fmaf(f[i], bf, cf);
                if (WHICH == 4) h[i] = fmaf(h[i], bf, cf);
            }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += a[i] + f[i] + h[i];
        if (s == 12345.678) sink_d[0] = s;
    } else {
        unsigned long long a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float2 f = make_float2(1.0f + threadIdx.x * 1e-6f + i, 2.0f + i);
            a[i] = *reinterpret_cast<unsigned long long *>(&f);
        }
        float2 bf = make_float2(0.999999f, 0.999998f), cf = make_float2(1e-7f, 2e-7f);
        const unsigned long long b = *reinterpret_cast<unsigned long long *>(&bf);
        const unsigned long long c = *reinterpret_cast<unsigned long long *>(&cf);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= a[i];
        if (s == 0x12345678ULL) sink_d[0] = (double)s;
    }
}


// Candidate-loop microbenchmarks: the inner loops of the two culled density kernels over an L1-resident run of
// candidates, no per-tile work -- what the loops themselves sustain (pair tests / s = lanes * 16 nodes).
template <int WHICH>
__global__ void __launch_bounds__(kDensThreads) k_loop_peak(int trips, const double2 *__restrict__ pts,
                                                            const uint2 *__restrict__ ptsq, double *sink) {
    const int lane = threadIdx.x & 31;
    if (WHICH == 0) {
        double acc[kTile][kTile];
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) acc[i][j] = 0.0;
        double2 pn = pts[lane];
        for (int t = 0; t < trips; ++t) {
            const double2 p = pn;
            pn = pts[((t + 1) & 31) * 32 + lane];
            accumulate_point(p, 0.25, 0.25, 33.0, 0.13, 0.26, 0.39, acc);
        }
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) s += acc[i][j];
        if (s == 12345.678) sink[0] = s;
    } else {
        float acc[kTile][kTile], hi[kTile][kTile];
        double acc64[kTile][kTile];
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) { acc[i][j] = 0.f; hi[i][j] = 0.f; acc64[i][j] = 0.0; }
        const TileOrigin32 org{0x10000000u, 0x10000000u, 0.25, 0.25, 33.0, 1.862645149230957e-09f};
#if KDEJS_FP32_FROM_DOUBLES
        const double2 *src = pts;
#else
        const uint2 *src = ptsq;
#endif
        Cand32 pn = src[lane];
        int budget = kFoldTrips;
        for (int t = 0; t < trips; ++t) {
            const Cand32 p = pn;
            pn = src[((t + 1) & 31) * 32 + lane];
            accumulate_point32(p, org, 0.13f, 0.26f, 0.39f, 1.0f, acc);
            if (WHICH == 2 && --budget == 0) { fold32(acc, hi); budget = kFoldTrips; }
        }
        fold32(acc, hi);
        flush32(hi, acc64);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kTile; ++i)
#pragma unroll
            for (int j = 0; j < kTile; ++j) s += acc64[i][j];
        if (s == 12345.678) sink[0] = s;
    }
}

}  // namespace kdejs
