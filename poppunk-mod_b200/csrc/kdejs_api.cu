// kdejs_api.cu -- C ABI (include/kdejs.h) over the sm_100a kernels in kdejs_kernels.cuh.
//
// Host-side structure: one lazily created context per device: a compute stream, a copy stream and ONE set of
// scratch buffers (the waves of a call are ordered on the compute stream anyway; the H2D copies of later waves
// run ahead on the copy stream into a ring of raw-input buffers).  Evaluations are grouped into waves of
// up to kMaxWaveItems; a wave's descriptors travel as a __grid_constant__ kernel parameter, so
// there is no descriptor staging buffer to race with asynchronous callers.  Entry points that enqueue on a
// caller's stream and return without synchronising (*_dev) leave an event behind that every later user of the
// scratch waits for (see Ctx::scratch_busy).
#include "../../include/kdejs.h"
#include "kdejs_kernels.cuh"
#include "kdejs_stager.h"
#include "kdejs_tsv.cuh"

#include <nvtx3/nvToolsExt.h>     // header-only NVTX 3: ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace kdejs;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char b_[512];                                                                     \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                     __FILE__, __LINE__);                                                     \
            return fail(KDEJS_ERR_CUDA, b_);                                                  \
        }                                                                                     \
    } while (0)
#define OKR(call) do { int r_ = (call); if (r_ != KDEJS_OK) return r_; } while (0)

// Host ranges of at least this size that are not page-locked go through the staging threads (kdejs_stager.h);
// smaller ones are not worth waking them for.
static const size_t kStageMinBytes = getenv("KDEJS_STAGE_MIN") ? (size_t)atoll(getenv("KDEJS_STAGE_MIN")) : ((size_t)4 << 20);

// Upload several host ranges in order on `st`.  Page-locked sources are copied directly; pageable ones are
// staged when together they are large enough.
int h2d_ranges(int device, const std::vector<Stager::Range> &ranges, cudaStream_t st) {
    std::vector<char> staged(ranges.size(), 0);
    std::vector<Stager::Range> plan;
    size_t pageable_bytes = 0;
    for (size_t i = 0; i < ranges.size(); ++i)
        if (ranges[i].bytes && Stager::pageable(ranges[i].src)) { staged[i] = 1; pageable_bytes += ranges[i].bytes; }
    Stager *sg = (pageable_bytes >= kStageMinBytes) ? Stager::get(device) : nullptr;
    if (sg) {
        for (size_t i = 0; i < ranges.size(); ++i) if (staged[i]) plan.push_back(ranges[i]);
        sg->plan(plan);
    }
    for (size_t i = 0; i < ranges.size(); ++i) {
        if (!ranges[i].bytes) continue;
        if (sg && staged[i]) CU(sg->issue_range(st));
        else CU(cudaMemcpyAsync(ranges[i].dst, ranges[i].src, ranges[i].bytes, cudaMemcpyHostToDevice, st));
    }
    return KDEJS_OK;
}

struct NvtxRange {     // RAII range on the calling thread (Nsight Systems / ncu --nvtx timelines: entry points and waves)
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KDEJS_OK;
        if (p) {
            // Work from ANY stream (a caller's stream of a *_dev entry point included) may still be using the old
            // block: drain the device before it is released.  Growth is rare (first calls, larger problems).
            CU(cudaDeviceSynchronize());
            CU(cudaFree(p));
        }
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return KDEJS_OK;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Slot {
    DevBuf scaled, sorted, sorted_q, cellid, blockhist, cellstart, moments, dens, pz, zout, tilesum;
    DevBuf rstat, rpart, tickets, jspart, dense_partial, worklist, redo, counters, totals, warpprefix, tileflag, misc;
    DevBuf partinfo, redo_partinfo, split_partial, split_ticket;     // very heavy tiles processed in parts (SplitBufs)
    DevBuf bandtot;                                                   // band mode: per-cell totals over ranks
};

struct Params {
    int R; double h, gamma, eps; int log_flag, kernel, algo;
    double gmin = 0.0, gmax = 1.0;
};

struct Pending { int kind; cudaEvent_t e0, e1; };

struct Ctx {
    int dev = -1;
    std::atomic<bool> ready{false};
    int sm_count = 0;
    int dens_blocks_per_sm = 1, dens32_blocks_per_sm = 1, minmax_blocks_per_sm = 4;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    // Two sets of scratch buffers (slot, tickets, counters, rstat).  Work on ONE stream always reuses the slot it
    // used last (stream order protects it); a second stream gets the other slot, so two callers' waves can overlap
    // on the GPU (the sort of one beside the density kernel of the other).  An entry point that returns with work
    // still enqueued records busy_ev[slot] on its stream; a user on a DIFFERENT stream that needs that slot waits
    // for it.  `cur` is the slot of the call in progress (callers hold `mu`).
    static constexpr int kSlots = 2;
    Slot slots[kSlots];
    int cur = 0;
    Slot &S() { return slots[cur]; }
    cudaEvent_t busy_ev[kSlots] = {nullptr, nullptr};
    cudaStream_t busy_stream[kSlots] = {nullptr, nullptr};
    bool busy_pending[kSlots] = {false, false};
    uint64_t last_use[kSlots] = {0, 0}, use_tick = 0;
    DevBuf pool_raw, js_out, status_out, sim_ring, summary;
    // tables parsed on the device (kdejs_discrepancy_batch_text): two text buffers (the copy stream fills one while
    // the other is parsed), chunk row counts and row totals per buffer, the parsed tables, per-file status
    DevBuf tsv_text[2], tsv_chunkrows[2], tsv_rows[2], tsv_sims, tsv_status;
    int64_t *h_tsv_rows = nullptr;        // pinned, 2 x kMaxTsvFiles
    std::vector<cudaEvent_t> free_plain_events;
    double *h_js = nullptr; int32_t *h_status = nullptr; size_t h_cap = 0;   // pinned results
    std::mutex mu;
    // measurement
    bool prof_on = false;
    int prof_mode = 0;                 // 1: every kernel, 2: the density kernel only
    double prof_ms[KDEJS_NUM_KERNELS] = {0};
    int64_t prof_n[KDEJS_NUM_KERNELS] = {0};
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> free_events;
    int64_t launches = 0;
    unsigned long long *d_work = nullptr, *d_stats32 = nullptr;
    // row-sharded evaluation state (phase 1 -> phase 2)
    bool shard_valid = false;
    WaveDesc shard_wd; GridGeom shard_geom; int shard_kb = 0, shard_ke = 0, shard_slot = 0;
    // band mode state (kdejs_band_phase1_dev -> kdejs_band_phase2_dev)
    bool band_valid = false;
    GridGeom band_geom; int band_ty0 = 0, band_ty1 = 0, band_slot = 0;
    // band mode over peer memory (kdejs_band_peer_*): this rank's locally sorted points live in peer_buf, exported
    // through a CUDA IPC handle; peer_map[r] is rank r's buffer mapped into this process
    DevBuf peer_buf;
    unsigned char peer_handle[64] = {0};
    bool peer_handle_valid = false;
    struct PeerMap { unsigned char handle[64]; void *ptr; bool open; };
    PeerMap peer_map[kMaxPeers] = {};
    int peer_world = 0, peer_rank = -1;
    std::vector<void *> peer_retired;
};

constexpr int kMaxDevices = 64;
Ctx g_ctx[kMaxDevices];
std::mutex g_init_mu;

int init_ctx(Ctx &c) {
    CU(cudaStreamCreateWithFlags(&c.s_compute, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c.s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < Ctx::kSlots; ++i) {
        CU(cudaEventCreateWithFlags(&c.busy_ev[i], cudaEventDisableTiming));
        OKR(c.slots[i].tickets.ensure(sizeof(unsigned) * 3 * kMaxWaveSets));   // self-resetting tickets
        CU(cudaMemset(c.slots[i].tickets.p, 0, c.slots[i].tickets.cap));
    }
    CU(cudaMalloc(&c.d_work, sizeof(unsigned long long)));
    CU(cudaMemset(c.d_work, 0, sizeof(unsigned long long)));
    CU(cudaMalloc(&c.d_stats32, 8 * sizeof(unsigned long long)));
    CU(cudaMemset(c.d_stats32, 0, 8 * sizeof(unsigned long long)));
    CU(cudaFuncSetAttribute(k_scale_count, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            kSortWarps * kMaxCells * (int)sizeof(uint16_t)));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.dens_blocks_per_sm, k_density_binned<false>, kDensThreads, 0));
    c.dens_blocks_per_sm = std::max(1, c.dens_blocks_per_sm);
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.dens32_blocks_per_sm, k_density_binned32<false>, kDensThreads, 0));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.minmax_blocks_per_sm, k_minmax, kMinmaxThreads, 0));
    c.minmax_blocks_per_sm = std::max(1, c.minmax_blocks_per_sm);
    c.dens32_blocks_per_sm = std::max(1, c.dens32_blocks_per_sm);
    // fewer resident density blocks leave registers for another stream's sort kernels (two-stream callers)
    if (const char *e = getenv("KDEJS_DENS_BLOCKS")) {
        const int v = atoi(e);
        if (v > 0) { c.dens_blocks_per_sm = std::min(c.dens_blocks_per_sm, v); c.dens32_blocks_per_sm = std::min(c.dens32_blocks_per_sm, v); }
    }
    return KDEJS_OK;
}

// Release everything a context owns (shutdown, or a failed init).  The caller holds g_init_mu.
void destroy_ctx(Ctx &c) {
    for (Slot &s : c.slots)
        for (DevBuf *b : {&s.scaled, &s.sorted, &s.sorted_q, &s.moments, &s.cellid, &s.blockhist, &s.cellstart, &s.dens,
                          &s.pz, &s.zout, &s.tilesum, &s.rstat, &s.rpart, &s.tickets, &s.jspart, &s.dense_partial,
                          &s.worklist, &s.redo, &s.counters, &s.totals, &s.warpprefix, &s.tileflag, &s.misc,
                          &s.partinfo, &s.redo_partinfo, &s.split_partial, &s.split_ticket, &s.bandtot}) {
            if (b->p) cudaFree(b->p);
            b->p = nullptr; b->cap = 0;
        }
    for (auto &m : c.peer_map) {
        if (m.open) cudaIpcCloseMemHandle(m.ptr);
        m.open = false; m.ptr = nullptr;
    }
    c.peer_world = 0; c.peer_rank = -1; c.peer_handle_valid = false;
    for (void *q : c.peer_retired) cudaFree(q);
    c.peer_retired.clear();
    if (c.h_tsv_rows) { cudaFreeHost(c.h_tsv_rows); c.h_tsv_rows = nullptr; }
    for (DevBuf *b : {&c.pool_raw, &c.js_out, &c.status_out, &c.sim_ring, &c.summary, &c.peer_buf, &c.tsv_text[0],
                      &c.tsv_text[1], &c.tsv_chunkrows[0], &c.tsv_chunkrows[1], &c.tsv_rows[0], &c.tsv_rows[1],
                      &c.tsv_sims, &c.tsv_status}) {
        if (b->p) cudaFree(b->p);
        b->p = nullptr; b->cap = 0;
    }
    for (auto e : c.free_plain_events) cudaEventDestroy(e);
    c.free_plain_events.clear();
    if (c.h_js) { cudaFreeHost(c.h_js); cudaFreeHost(c.h_status); c.h_js = nullptr; c.h_status = nullptr; c.h_cap = 0; }
    for (auto &pd : c.pending) { cudaEventDestroy(pd.e0); cudaEventDestroy(pd.e1); }
    c.pending.clear();
    for (auto e : c.free_events) cudaEventDestroy(e);
    c.free_events.clear();
    if (c.d_work) { cudaFree(c.d_work); c.d_work = nullptr; }
    if (c.d_stats32) { cudaFree(c.d_stats32); c.d_stats32 = nullptr; }
    for (int i = 0; i < Ctx::kSlots; ++i) {
        if (c.busy_ev[i]) { cudaEventDestroy(c.busy_ev[i]); c.busy_ev[i] = nullptr; }
        c.busy_pending[i] = false;
        c.busy_stream[i] = nullptr;
    }
    if (c.s_compute) { cudaStreamDestroy(c.s_compute); c.s_compute = nullptr; }
    if (c.s_copy) { cudaStreamDestroy(c.s_copy); c.s_copy = nullptr; }
    c.shard_valid = false;
    c.band_valid = false;
}

// Every user of the scratch calls this first with the stream it is about to enqueue on: picks the slot (c.cur).
// want >= 0: the caller continues work whose state lives in that slot (phase 2 after phase 1).
int acquire_scratch(Ctx &c, cudaStream_t st, int want = -1) {
    for (int i = 0; i < Ctx::kSlots; ++i)
        if (c.busy_pending[i] && cudaEventQuery(c.busy_ev[i]) == cudaSuccess) c.busy_pending[i] = false;
    int pick = want;
    if (pick < 0)
        for (int i = 0; i < Ctx::kSlots; ++i)
            if (c.busy_stream[i] == st && (c.busy_pending[i] || c.last_use[i])) { pick = i; break; }   // this stream's slot
    if (pick < 0)
        for (int i = 0; i < Ctx::kSlots; ++i)
            if (!c.busy_pending[i] && !c.last_use[i]) { pick = i; break; }                            // never used
    if (pick < 0) {                                                                                   // least recently used
        pick = 0;
        for (int i = 1; i < Ctx::kSlots; ++i)
            if (c.last_use[i] < c.last_use[pick]) pick = i;
    }
    if (c.busy_pending[pick] && c.busy_stream[pick] != st) CU(cudaStreamWaitEvent(st, c.busy_ev[pick], 0));
    c.cur = pick;
    c.last_use[pick] = ++c.use_tick;
    c.busy_stream[pick] = st;
    return KDEJS_OK;
}
// ... and this last: `synced` = the entry point synchronised `st` before returning (nothing is left in flight).
int release_scratch(Ctx &c, cudaStream_t st, bool synced) {
    if (synced) {
        // a wait on busy_ev was enqueued before our work, so whatever recorded it has finished too
        c.busy_pending[c.cur] = false;
        return KDEJS_OK;
    }
    CU(cudaEventRecord(c.busy_ev[c.cur], st));
    c.busy_pending[c.cur] = true;
    c.busy_stream[c.cur] = st;
    return KDEJS_OK;
}

int get_ctx(int device, Ctx **out) {
    if (device < 0 || device >= kMaxDevices) return fail(KDEJS_ERR_ARG, "device index out of range");
    Ctx &c = g_ctx[device];
    if (!c.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_init_mu);
        if (!c.ready.load(std::memory_order_relaxed)) {
            int n = 0;
            CU(cudaGetDeviceCount(&n));
            if (device >= n) return fail(KDEJS_ERR_CUDA, "no such CUDA device (kdejs has no CPU fallback)");
            CU(cudaSetDevice(device));
            cudaDeviceProp prop;
            CU(cudaGetDeviceProperties(&prop, device));
            if (prop.major < 10)
                return fail(KDEJS_ERR_CUDA, std::string("kdejs is built for sm_100a only; device is ") + prop.name);
            c.dev = device;
            c.sm_count = prop.multiProcessorCount;
            const int rc = init_ctx(c);
            if (rc != KDEJS_OK) { destroy_ctx(c); return rc; }       // a failed init leaves nothing behind for the retry
            c.ready.store(true, std::memory_order_release);
        }
    }
    CU(cudaSetDevice(device));
    *out = &c;
    return KDEJS_OK;
}

struct KTimer {       // CUDA-event bracket around one launch on the launching stream
    Ctx &c; int kind; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool active = false;
    KTimer(Ctx &c_, int kind_, cudaStream_t s_) : c(c_), kind(kind_), s(s_) {
        active = c.prof_on && (c.prof_mode == 1 || kind == 4);
        if (!active) return;
        for (cudaEvent_t *e : {&e0, &e1}) {
            if (!c.free_events.empty()) { *e = c.free_events.back(); c.free_events.pop_back(); }
            else cudaEventCreate(e);
        }
        cudaEventRecord(e0, s);
    }
    ~KTimer() {
        c.launches++;
        if (!active) return;
        cudaEventRecord(e1, s);
        c.pending.push_back({kind, e0, e1});
    }
};

int resolve_pending(Ctx &c) {
    if (c.pending.empty()) return KDEJS_OK;
    CU(cudaDeviceSynchronize());
    for (auto &p : c.pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) { c.prof_ms[p.kind] += ms; c.prof_n[p.kind]++; }
        c.free_events.push_back(p.e0);
        c.free_events.push_back(p.e1);
    }
    c.pending.clear();
    return KDEJS_OK;
}

int check_params(const Params &p) {
    if (p.R < 1 || p.R > 16384) return fail(KDEJS_ERR_ARG, "resolution must be in [1, 16384]");
    if (!(p.h > 0.0) || !std::isfinite(p.h)) return fail(KDEJS_ERR_ARG, "bandwidth must be positive and finite");
    if (!(p.gmax > p.gmin)) return fail(KDEJS_ERR_ARG, "grid maximum must exceed grid minimum");
    if (p.kernel != KDEJS_KERNEL_EPANECHNIKOV && p.kernel != KDEJS_KERNEL_GAUSSIAN)
        return fail(KDEJS_ERR_ARG, "unknown kernel");
    if (p.algo != KDEJS_ALGO_BINNED && p.algo != KDEJS_ALGO_DENSE && p.algo != KDEJS_ALGO_DENSE64 &&
        p.algo != KDEJS_ALGO_BINNED32)
        return fail(KDEJS_ERR_ARG, "unknown algorithm");
    if (p.kernel == KDEJS_KERNEL_GAUSSIAN && (p.algo == KDEJS_ALGO_BINNED || p.algo == KDEJS_ALGO_BINNED32))
        return fail(KDEJS_ERR_ARG, "the gaussian kernel has unbounded support: use a dense algorithm");
    if (std::isnan(p.gamma) || std::isnan(p.eps)) return fail(KDEJS_ERR_ARG, "gamma/eps must not be NaN");
    return KDEJS_OK;
}

GridGeom make_geom(const Params &p) {
    GridGeom g;
    const double ext = p.gmax - p.gmin;
    g.R = p.R;
    g.NT = (p.R + kTile - 1) / kTile;
    g.gmin = p.gmin;
    g.step = (p.R > 1) ? ext / (double)(p.R - 1) : 0.0;   // numpy.linspace step
    g.h = p.h;
    g.inv_h = 1.0 / p.h;
    g.dh = g.step * g.inv_h;
    g.margin = 1e-9 * ext;
    // culling cells: side >= h/2 and at most 64 per side; on fine grids (a cell would span more than two tiles)
    // side >= h/4 and up to 128 per side -- the sort's shared-memory histograms grow 4x, which only pays when the
    // density kernel dominates; always at most kMaxRows cell rows per tile
    double reach = (kTile - 1) * g.step + 2.0 * p.h + 2.0 * g.margin;
    int C = std::min((int)std::floor(2.0 * ext / p.h), 64);
    if (ext / std::max(C, 1) > 2.0 * kTile * g.step && g.step > 0.0)
        C = std::min((int)std::floor(4.0 * ext / p.h), kMaxCellsSide);
    if (const char *e = getenv("KDEJS_CELLS")) C = atoi(e);
    C = std::min(C, (int)std::floor((kMaxRows - 2) * ext / reach));
    C = std::max(1, std::min(C, kMaxCellsSide));
    g.C = C;
    g.cell_x0 = p.gmin;
    g.cell_invw = (double)C / ext;
    g.cell_w = ext / (double)C;
    // sklearn _log_kernel_norm, d = 2 (sklearn:neighbors/_binary_tree.pxi.tp:448-476)
    const double LOG_PI = 1.1447298858494001741, LOG_2PI = 1.8378770664093454836;
    double factor = (p.kernel == KDEJS_KERNEL_GAUSSIAN) ? LOG_2PI : (LOG_PI - std::lgamma(2.0)) + std::log(0.5);
    g.knorm = std::exp(-factor - 2.0 * std::log(p.h));
    g.gamma = p.gamma;
    g.eps = p.eps;
    g.gamma_mode = (p.gamma == 1.0) ? 1 : (p.gamma == 0.5) ? 2 : (p.gamma == 0.25) ? 3 : 0;
    g.tile_row_begin = 0;
    g.tile_row_end = g.NT;
    g.tile_col_begin = 0;
    g.tile_col_end = g.NT;
    g.sets_per_group = kMaxWaveSets;      // one group until run_wave() knows the wave
    {
        int split_log2 = kSplitLog2, part_log2 = kPartLog2;
        if (const char *e = getenv("KDEJS_SPLIT_LOG2")) split_log2 = std::max(8, std::min(30, atoi(e)));
        if (const char *e = getenv("KDEJS_PART_LOG2")) part_log2 = std::max(8, std::min(split_log2, atoi(e)));
        g.split_cfg = split_log2 | (part_log2 << 8);
    }
    g.heavy_log2 = 13;
    if (const char *e = getenv("KDEJS_HEAVY_LOG2")) g.heavy_log2 = std::max(6, std::min(30, atoi(e)));
    // Cells that lie wholly inside the support of every node of a tile enter through their moments.  A cell
    // concentric with a tile qualifies up to a centre offset rho; worth its bookkeeping once that region is at
    // least about two cells across (fine grids: R >= ~512 at h = 0.03; never at the headline R = 256).
    {
        const double half = 0.5 * ((kTile - 1) * g.step + g.cell_w);
        const double rho = (p.h > half) ? std::sqrt(p.h * p.h - half * half) - half : -1.0;
        // 1: eligible, run_wave() decides from the number of points per cell; 2: forced (KDEJS_MOMENTS=1)
        g.use_moments = (p.kernel == KDEJS_KERNEL_EPANECHNIKOV && rho >= g.cell_w && C >= 4) ? 1 : 0;
        if (const char *e = getenv("KDEJS_MOMENTS")) g.use_moments = (atoi(e) != 0 && p.kernel == KDEJS_KERNEL_EPANECHNIKOV && C >= 4) ? 2 : 0;
    }
    // fp32 path: sorted points as 32-bit fixed point in units of h * 2^-q_shift, modulo 2^32.  The largest
    // coordinate difference between a tile's first node and one of its candidates (clamped to 2h beyond the grid)
    // must stay below 2^31 units: left of the node h + one cell (+ 2h when clamped), right of it tile side + h + cell.
    {
        const double cw_h = g.cell_w / p.h, tile_h = (kTile - 1) * g.dh;
        const double maxdiff_h = std::max(3.0 + cw_h, tile_h + 1.0 + cw_h) + 0.01;
        int shift = (int)std::floor(std::log2(2147483648.0 / maxdiff_h));
        shift = std::min(shift, 29);                  // the padding candidate sits 2^30 units (>= 2 h) left of the node
        g.q_shift = (shift >= 26 && p.kernel == KDEJS_KERNEL_EPANECHNIKOV && p.R > 1) ? shift : 0;
        g.q_pad = 0;
        g.q_scale = std::ldexp(g.inv_h, shift);
        g.q_lo = p.gmin - 2.0 * p.h;
        g.q_hi = p.gmax + 2.0 * p.h;
        g.q_unit = (float)std::ldexp(1.0, -shift);
        g.q_dh = (float)g.dh;
    }
    return g;
}

// What one wave needs beyond the descriptors.
struct WaveIO {
    double *out_js = nullptr;        // device, indexed item_base + item
    int32_t *out_status = nullptr;   // device
    int item_base = 0;
    bool want_z = false;             // fill slot.zout [set][G]
    bool density_only = false;       // stop after the density kernel (kde_density)
    bool identity_scaler = false;
    bool shard = false;              // row-sharded: no k_js, caller runs the phases
    int balance_rank = -1, balance_world = 0;   // row-sharded with work-balanced bands chosen here
    int *out_tile_rows = nullptr;               // [2]: the tile rows this rank ended up with
    // band mode (one evaluation over several GPUs, kdejs_band_*)
    bool rstat_ready = false;        // slot.rstat already holds the (global) extrema: no k_minmax
    bool sort_only = false;          // stop after k_place (the caller ships the sorted points)
    bool presorted = false;          // slot.sorted / sorted_q / cellstart were assembled by k_band_merge: no sort
    int force_moments = -1;          // -1: decide from this wave's point count; 0 / 1: decided by the caller (all ranks alike)
    const int64_t *n_global = nullptr;   // [n_sets]: sizes of the WHOLE sets (the 1/N of the density) when this wave holds a part
};

constexpr int64_t kMinPoints32 = 64;     // sets smaller than this are summed in fp64 whatever the algorithm says

int sort_shape(const WaveDesc &wd, int sm_count, int ncells, int &chw, int &nb_max) {
    // Points per warp sub-chunk.  Every sort block pays a fixed cost for its C*C counters (zeroing, prefix,
    // histogram write-out), so chunks grow with the wave until the launch still has ~4 blocks per SM
    // (8 per SM was measured: k_scale_count -4 %, k_scan +75 %).
    int nmax = 0;
    int64_t total = 0;
    for (int s = 0; s < wd.n_sets; ++s) { nmax = std::max(nmax, wd.set[s].n); total += wd.set[s].n; }
    // ... rounded DOWN to a power of two for the 64 x 64 cell geometry (rounding up left 3.5 blocks per SM with a
    // nearly empty last block per set: k_scale_count 2.06 -> 2.84 us per evaluation).
    // With the 128 x 128 cell geometry a block's counters fill half an SM's shared memory (one block per SM is
    // resident anyway) and every block costs k_scan another pass over 16 384 cells: exactly one block per SM there,
    // chunks cut to measure (a band of config C5 on eight GPUs: 2 x 77 power-of-two blocks were 1.04 waves of the
    // 148 SMs and took twice the time of 2 x 74).
    if (ncells > 4096) {
        const int blocks_per_set = std::max(1, sm_count / std::max(1, (int)wd.n_sets));
        chw = 512;
        for (int s = 0; s < wd.n_sets; ++s) {
            const int64_t per_warp = ((int64_t)wd.set[s].n + (int64_t)blocks_per_set * kSortWarps - 1) / ((int64_t)blocks_per_set * kSortWarps);
            chw = (int)std::max<int64_t>(chw, std::min<int64_t>((per_warp + 31) / 32 * 32, kMaxChw));
        }
    } else {
        int64_t per_warp = total / ((int64_t)4 * sm_count * kSortWarps);
        chw = 512;
        while (2 * chw <= per_warp && chw < kMaxChw) chw <<= 1;
    }
    nb_max = std::max(1, (int)(((int64_t)nmax + (int64_t)chw * kSortWarps - 1) / ((int64_t)chw * kSortWarps)));
    return KDEJS_OK;
}

// Enqueue the launches of one wave on `st`.  wd.set[].nb is filled here.
int run_wave(Ctx &c, Slot &sl, cudaStream_t st, const Params &p, const GridGeom &g_in, WaveDesc &wd,
             const WaveIO &io) {
    NvtxRange nvtx_wave(io.sort_only ? "kdejs wave: sort only" : io.presorted ? "kdejs wave: band density" : "kdejs wave");
    GridGeom g = g_in;
    const int n_sets = wd.n_sets, n_items = n_sets / 2;
    const size_t G = (size_t)p.R * p.R;
    int64_t P = 0;
    for (int s = 0; s < n_sets; ++s) {
        wd.set[s].off = P;
        P += wd.set[s].n;
        wd.set[s].knorm_over_n = g.knorm / (double)(io.n_global ? io.n_global[s] : (int64_t)wd.set[s].n);
    }
    int chw, nb_max;
    sort_shape(wd, c.sm_count, g.C * g.C, chw, nb_max);
    for (int s = 0; s < n_sets; ++s)
        wd.set[s].nb = (int)(((int64_t)wd.set[s].n + (int64_t)chw * kSortWarps - 1) / ((int64_t)chw * kSortWarps));
    const int ncells = g.C * g.C;
    const bool binned = (p.algo == KDEJS_ALGO_BINNED || p.algo == KDEJS_ALGO_BINNED32);
    // the fp32 kernel needs a geometry its fixed-point format covers; otherwise the fp64 kernel does the work
    const int force32 = getenv("KDEJS_FP32") ? atoi(getenv("KDEJS_FP32")) : -1;     // 0 / 1: override the algorithm's choice
    // ... and sets large enough for the absolute part of its error bound (1e-7 of the LARGEST node sum) to mean
    // something: with a handful of points no node sum exceeds ~1 and fp32 terms are themselves 1e-7 off
    int64_t min_n = INT64_MAX;
    for (int s = 0; s < wd.n_sets; ++s) min_n = std::min<int64_t>(min_n, io.n_global ? io.n_global[s] : (int64_t)wd.set[s].n);
    bool fast32 = binned && g.q_shift > 0 && min_n >= kMinPoints32 &&
                  (force32 >= 0 ? force32 != 0 : p.algo == KDEJS_ALGO_BINNED32);
    // a moment cell costs a tile as much as one point plus some per-row bookkeeping: it pays once cells hold
    // hundreds of points (measured: -40 % density time at 1200 points per cell, +48 % at 49)
    if (io.force_moments >= 0 && g_in.use_moments == 1) g.use_moments = io.force_moments ? 1 : 0;
    else if (g.use_moments == 1 && P < (int64_t)128 * n_sets * ncells) g.use_moments = 0;
    // Fine grids with dense cells (config C5) stay on the fp64 kernel unless fp32 is forced: most of their sum
    // comes from moment cells (fp64 either way), and their very heavy edge tiles hold nodes whose sum stays below
    // the fp32 certification threshold (bias x candidates), so a third of the work was redone in fp64 anyway
    // (measured, 10 M + 10 M points: 7.2 ms + 4.2 ms redo against 12.8 ms fp64, and a band-to-band imbalance).
    if (g.use_moments && force32 <= 0) fast32 = false;
    // Plain discrepancies (no grids requested, one launch for the whole grid): tiles without candidates are only
    // flagged and the divergence kernel reads the flag instead of 16 node values -- ~70 % of the grid is never
    // written or read.
    const bool lazy_empty = binned && !io.want_z && !io.density_only && !io.shard && io.balance_world == 0 &&
                            g.tile_row_begin == 0 && g.tile_row_end == g.NT && g.tile_col_begin == 0 &&
                            g.tile_col_end == g.NT;
    if (lazy_empty) OKR(sl.tileflag.ensure((size_t)n_sets * g.NT * g.NT));
    const int nparts = binned ? g.NT * g.NT : (int)((G + 255) / 256);

    if (!binned) OKR(sl.scaled.ensure((size_t)P * sizeof(double2)));
    OKR(sl.dens.ensure((size_t)n_sets * G * sizeof(double)));
    OKR(sl.pz.ensure((size_t)n_sets * G * sizeof(double)));
    OKR(sl.tilesum.ensure((size_t)n_sets * nparts * sizeof(double)));
    OKR(sl.rstat.ensure(sizeof(RawStat) * kMaxWaveSets));
    // blocks per raw set of the min/max pass: ONE wave of resident blocks over the whole launch (18 per set at 65 raw
    // sets were 1.58 waves of the 740 resident slots: 0.53 -> 0.48 us per evaluation with 11), few enough that a
    // block's reduction and ticket do not dominate its few hundred points
    static const int mm_env = getenv("KDEJS_MM_BLOCKS") ? atoi(getenv("KDEJS_MM_BLOCKS")) : 0;
    const int mm_blocks = mm_env > 0 ? std::min(mm_env, 64)
                                     : std::max(8, std::min(64, (c.minmax_blocks_per_sm * c.sm_count) / std::max(1, (int)wd.n_raw)));
    OKR(sl.rpart.ensure(sizeof(RawStat) * kMaxWaveSets * mm_blocks));
    const int js_blocks = (int)((G + kJsNodesPerBlock - 1) / kJsNodesPerBlock);
    OKR(sl.jspart.ensure(sizeof(double) * 2 * (size_t)std::max(1, n_items) * std::max(js_blocks, g.NT)));
    if (io.want_z) OKR(sl.zout.ensure((size_t)n_sets * G * sizeof(double)));
    if (binned) {
        OKR(sl.sorted.ensure((size_t)P * sizeof(double2)));
        OKR(sl.cellid.ensure((size_t)P * sizeof(uint32_t)));        // per point: cell id | rank << 16
        OKR(sl.warpprefix.ensure((size_t)n_sets * nb_max * kSortWarps * ncells * sizeof(uint16_t)));
        OKR(sl.blockhist.ensure((size_t)n_sets * nb_max * ncells * sizeof(uint32_t)));
        OKR(sl.cellstart.ensure((size_t)n_sets * (ncells + 1) * sizeof(uint32_t)));
        if (g.use_moments) OKR(sl.moments.ensure((size_t)n_sets * ncells * kMomentDoubles * sizeof(double)));
        // set groups of the work list: a group's sorted points (16 B each) should fit the L2 a few times over
        {
            const double group_bytes = 40.0e6;
            int n_groups = (int)std::ceil((double)P * 16.0 / group_bytes);
            n_groups = std::max(1, std::min(std::min(n_groups, kMaxGroups), n_sets));
            if (const char *e = getenv("KDEJS_GROUPS")) n_groups = std::max(1, std::min(std::min(atoi(e), kMaxGroups), n_sets));
            g.sets_per_group = (n_sets + n_groups - 1) / n_groups;
        }
        OKR(sl.counters.ensure(kAllCounters * sizeof(unsigned)));
        if (fast32 && !kFp32FromDoubles) OKR(sl.sorted_q.ensure((size_t)P * sizeof(uint2)));
    }

    unsigned *tk_minmax = sl.tickets.as<unsigned>();
    unsigned *tk_js = sl.tickets.as<unsigned>() + kMaxWaveSets;
    unsigned *tk_scan = sl.tickets.as<unsigned>() + 2 * kMaxWaveSets;

    if (!io.identity_scaler && !io.rstat_ready) {
        KTimer t(c, 0, st);
        k_minmax<<<dim3(mm_blocks, wd.n_raw), kMinmaxThreads, 0, st>>>(
            wd, p.log_flag, p.eps, sl.rpart.as<RawStat>(), tk_minmax, sl.rstat.as<RawStat>());
    }
    if (binned) {
        const size_t smem = (size_t)kSortWarps * ncells * sizeof(uint16_t);
        // one block per set scans <= 4096 cells over a few sort blocks fastest; otherwise one thread per cell
        // (measured per wave, tools/wave_probe.py: 13 sort blocks per set 23 vs 25 us, 25 blocks 39 vs 21 us)
        static const int force_wide = getenv("KDEJS_SCAN_WIDE") ? atoi(getenv("KDEJS_SCAN_WIDE")) : -1;
        const bool wide_scan = force_wide >= 0 ? force_wide != 0 : (ncells > 4096 || nb_max > 16);
        if (!io.presorted) {
            KTimer t(c, 1, st);
            k_scale_count<<<dim3(nb_max, n_sets), kSortThreads, smem, st>>>(
                wd, sl.rstat.as<RawStat>(), io.identity_scaler ? 0 : p.log_flag, p.eps,
                io.identity_scaler ? 1 : 0, g.C, g.cell_x0, g.cell_invw, chw, nb_max,
                nullptr, sl.cellid.as<uint32_t>(), sl.warpprefix.as<uint16_t>(),
                sl.blockhist.as<uint32_t>());
        }
        if (!io.presorted) {
            KTimer t(c, 2, st);
            if (wide_scan)
                k_scan_wide<<<dim3((ncells + 1023) / 1024, n_sets), 1024, 0, st>>>(
                    wd, ncells, nb_max, sl.blockhist.as<uint32_t>(), sl.cellstart.as<uint32_t>(),
                    sl.counters.as<unsigned>(), tk_scan);
            else
                k_scan_narrow<<<n_sets, 1024, 0, st>>>(wd, ncells, nb_max, sl.blockhist.as<uint32_t>(),
                                                       sl.cellstart.as<uint32_t>(), sl.counters.as<unsigned>());
        }
        if (!io.presorted) {
            KTimer t(c, 3, st);
            int nmax = 0;
            for (int s = 0; s < n_sets; ++s) nmax = std::max(nmax, wd.set[s].n);
            k_place<<<dim3((nmax + 255) / 256, n_sets), 256, 0, st>>>(
                wd, sl.rstat.as<RawStat>(), io.identity_scaler ? 0 : p.log_flag, p.eps,
                io.identity_scaler ? 1 : 0, ncells, chw, nb_max, sl.cellid.as<uint32_t>(),
                sl.warpprefix.as<uint16_t>(), sl.blockhist.as<uint32_t>(),
                wide_scan ? sl.cellstart.as<uint32_t>() : nullptr, sl.sorted.as<double2>(),
                fast32 && !kFp32FromDoubles ? sl.sorted_q.as<uint2>() : nullptr, g);
        }
        if (io.sort_only) { CU(cudaGetLastError()); return KDEJS_OK; }
        if (io.presorted) CU(cudaMemsetAsync(sl.counters.p, 0, kAllCounters * sizeof(unsigned), st));   // k_scan's job otherwise
        if (g.use_moments) {
            KTimer t(c, 7, st);
            k_cell_moments<<<dim3(ncells, n_sets), kMomentThreads, 0, st>>>(
                wd, g, sl.sorted.as<double2>(), sl.cellstart.as<uint32_t>(), sl.moments.as<double>());
        }
        if (io.balance_world > 0) {
            // cut the tile rows into bands of equal estimated work; all ranks derive the same cuts
            OKR(c.S().misc.ensure(64 + sizeof(double) * (size_t)g.NT));
            double *d_rw = c.S().misc.as<double>() + 8;
            {
                KTimer t(c, 6, st);
                k_tile_rowwork<<<g.NT, 128, 0, st>>>(wd, g, sl.cellstart.as<uint32_t>(), d_rw);
            }
            std::vector<double> rw(g.NT);
            CU(cudaMemcpyAsync(rw.data(), d_rw, sizeof(double) * (size_t)g.NT, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            double total = 0.0;
            for (double v : rw) total += v;
            std::vector<int> cut(io.balance_world + 1, g.NT);
            cut[0] = 0;
            double run = 0.0;
            int r = 1;
            for (int t = 0; t < g.NT && r < io.balance_world; ++t) {
                run += rw[t];
                while (r < io.balance_world && run >= total * r / io.balance_world) cut[r++] = t + 1;
            }
            g.tile_row_begin = cut[io.balance_rank];
            g.tile_row_end = cut[io.balance_rank + 1];
        }
        if (io.out_tile_rows) { io.out_tile_rows[0] = g.tile_row_begin; io.out_tile_rows[1] = g.tile_row_end; }
        const int tiles = (g.tile_row_end - g.tile_row_begin) * (g.tile_col_end - g.tile_col_begin);
        // Part slots for very heavy tiles (SplitBufs).  A tile is split into <= kMaxParts parts of ~2^kPartLog2
        // candidates; the sum of all tiles' candidates is at most (points) x (tiles one point can reach).
        const double reach = std::min((double)g.NT, (2.0 * p.h + g.cell_w) / (kTile * g.step > 0.0 ? kTile * g.step : 1.0) + 2.0);
        const double parts_bound = std::min((double)kMaxParts * tiles * n_sets,
                                            1.5 * (double)P * reach * reach / (double)(1 << (g.split_cfg >> 8)) + 64.0);
        const unsigned parts_cap = (unsigned)std::min(parts_bound, 8.0e6);
        const unsigned capacity = (unsigned)tiles * (unsigned)g.sets_per_group +      // entries per (group, class) list
                                  std::min(parts_cap, (unsigned)(kMaxParts - 1) * (unsigned)tiles * (unsigned)g.sets_per_group);
        OKR(sl.worklist.ensure((size_t)kNumLists * capacity * sizeof(uint32_t)));
        OKR(sl.partinfo.ensure((size_t)kMaxGroups * capacity * sizeof(uint2)));
        OKR(sl.split_partial.ensure((size_t)parts_cap * 16 * sizeof(double)));
        if ((size_t)parts_cap * sizeof(unsigned) > sl.split_ticket.cap) {
            OKR(sl.split_ticket.ensure((size_t)parts_cap * sizeof(unsigned)));
            CU(cudaMemsetAsync(sl.split_ticket.p, 0, sl.split_ticket.cap, st));     // self-resetting afterwards
        }
        const SplitBufs sb{sl.partinfo.as<uint2>(), sl.split_partial.as<double>(), sl.split_ticket.as<unsigned>(), parts_cap};
        if (fast32) {
            OKR(sl.redo.ensure((size_t)kNumLists * capacity * sizeof(uint32_t)));
            OKR(sl.redo_partinfo.ensure((size_t)kMaxGroups * capacity * sizeof(uint2)));
        }
        // the redo list of the fp32 path: a split tile comes back in the same parts, with the same part slots
        const SplitBufs sb_redo{sl.redo_partinfo.as<uint2>(), sl.split_partial.as<double>(), sl.split_ticket.as<unsigned>(), parts_cap};
        if (tiles > 0) {
            KTimer t(c, 6, st);
            k_tile_classify<<<dim3((tiles + 31) / 32, n_sets), 256, 0, st>>>(
                wd, g, sl.cellstart.as<uint32_t>(), sl.dens.as<double>(), sl.pz.as<double>(),
                sl.tilesum.as<double>(), sl.worklist.as<uint32_t>(), sl.counters.as<unsigned>(), capacity,
                c.prof_on ? c.d_work : nullptr, lazy_empty ? sl.tileflag.as<uint8_t>() : nullptr, sb);
        }
        if (tiles > 0 && fast32) {
            // nodes whose fp32 sum stays below this are not trusted to fp32 (relative error ~1.5e-7 / sum): their
            // tile is walked a second time and, unless every such node is certified exactly zero, redone in fp64
            static const float fp32_min_thr = getenv("KDEJS_FP32_THR") ? (float)atof(getenv("KDEJS_FP32_THR")) : 0.0f;
            {
                KTimer t(c, 4, st);
                const int blocks = (int)std::min<int64_t>((int64_t)tiles * n_sets, (int64_t)c.sm_count * c.dens32_blocks_per_sm);
                if (g.use_moments)
                    k_density_binned32<true><<<blocks, kDensThreads, 0, st>>>(
                        wd, g, kFp32FromDoubles ? sl.sorted.as<Cand32>() : sl.sorted_q.as<Cand32>(), sl.cellstart.as<uint32_t>(), sl.moments.as<double>(),
                        sl.dens.as<double>(), sl.pz.as<double>(), sl.tilesum.as<double>(),
                        sl.worklist.as<uint32_t>(), sl.counters.as<unsigned>(), capacity, sl.redo.as<uint32_t>(),
                        fp32_min_thr, c.prof_on ? c.d_stats32 : nullptr, sb, sl.redo_partinfo.as<uint2>());
                else
                    k_density_binned32<false><<<blocks, kDensThreads, 0, st>>>(
                        wd, g, kFp32FromDoubles ? sl.sorted.as<Cand32>() : sl.sorted_q.as<Cand32>(), sl.cellstart.as<uint32_t>(), nullptr,
                        sl.dens.as<double>(), sl.pz.as<double>(), sl.tilesum.as<double>(),
                        sl.worklist.as<uint32_t>(), sl.counters.as<unsigned>(), capacity, sl.redo.as<uint32_t>(),
                        fp32_min_thr, c.prof_on ? c.d_stats32 : nullptr, sb, sl.redo_partinfo.as<uint2>());
            }
            // tiles with a node the fp32 arithmetic could not certify (a nearest candidate within ~1e-6 of the
            // support's edge): the fp64 kernel on the redo list -- usually empty, a few tiles per evaluation at most
            KTimer t(c, 8, st);
            const int blocks = (int)std::min<int64_t>((int64_t)tiles * n_sets, (int64_t)c.sm_count * c.dens_blocks_per_sm);
            if (g.use_moments)
                k_density_binned<true><<<blocks, kDensThreads, 0, st>>>(
                    wd, g, sl.sorted.as<double2>(), sl.cellstart.as<uint32_t>(), sl.moments.as<double>(),
                    sl.dens.as<double>(), sl.pz.as<double>(), sl.tilesum.as<double>(),
                    sl.redo.as<uint32_t>(), sl.counters.as<unsigned>() + kRedoBase, capacity, sb_redo);
            else
                k_density_binned<false><<<blocks, kDensThreads, 0, st>>>(
                    wd, g, sl.sorted.as<double2>(), sl.cellstart.as<uint32_t>(), nullptr,
                    sl.dens.as<double>(), sl.pz.as<double>(), sl.tilesum.as<double>(),
                    sl.redo.as<uint32_t>(), sl.counters.as<unsigned>() + kRedoBase, capacity, sb_redo);
        } else if (tiles > 0) {
            KTimer t(c, 4, st);
            const int blocks = (int)std::min<int64_t>((int64_t)tiles * n_sets, (int64_t)c.sm_count * c.dens_blocks_per_sm);
            if (g.use_moments)
                k_density_binned<true><<<blocks, kDensThreads, 0, st>>>(
                    wd, g, sl.sorted.as<double2>(), sl.cellstart.as<uint32_t>(), sl.moments.as<double>(),
                    sl.dens.as<double>(), sl.pz.as<double>(), sl.tilesum.as<double>(),
                    sl.worklist.as<uint32_t>(), sl.counters.as<unsigned>(), capacity, sb);
            else
                k_density_binned<false><<<blocks, kDensThreads, 0, st>>>(
                    wd, g, sl.sorted.as<double2>(), sl.cellstart.as<uint32_t>(), nullptr,
                    sl.dens.as<double>(), sl.pz.as<double>(), sl.tilesum.as<double>(),
                    sl.worklist.as<uint32_t>(), sl.counters.as<unsigned>(), capacity, sb);
        }
    } else {
        // dense fp64: scale only (C = 1: every point lands in cell 0, the histogram is unused)
        {
            KTimer t(c, 1, st);
            OKR(sl.blockhist.ensure((size_t)n_sets * nb_max * sizeof(uint32_t)));
            k_scale_count<<<dim3(nb_max, n_sets), kSortThreads, kSortWarps * sizeof(uint16_t), st>>>(
                wd, sl.rstat.as<RawStat>(), io.identity_scaler ? 0 : p.log_flag, p.eps,
                io.identity_scaler ? 1 : 0, 1, g.cell_x0, g.cell_invw, chw, nb_max,
                sl.scaled.as<double2>(), nullptr, nullptr, sl.blockhist.as<uint32_t>());
        }
        int nslices;
        if (p.algo == KDEJS_ALGO_DENSE) {
            int nmax = 0;
            for (int s = 0; s < n_sets; ++s) nmax = std::max(nmax, wd.set[s].n);
            nslices = (nmax + kD32Slice - 1) / kD32Slice;
            // one fp64 partial grid per 1024-point slice and set: the brute-force kernel is a small-problem tool
            if (nslices > 65535 || (double)nslices * n_sets * (double)G * 8.0 > 16.0e9)
                return fail(KDEJS_ERR_ARG, "algo=dense: point sets too large for the brute-force kernel (its slice "
                                           "partials would need " + std::to_string((size_t)((double)nslices * n_sets * G * 8.0 / 1e9)) +
                                           " GB); use the binned algorithm");
            OKR(sl.dense_partial.ensure((size_t)nslices * n_sets * G * sizeof(double)));
            KTimer t(c, 4, st);
            const int nbx = (p.R + kD32NodesX - 1) / kD32NodesX, nby = (p.R + kD32NodesY - 1) / kD32NodesY;
            dim3 grid(nbx * nby, n_sets, nslices);
            if (p.kernel == KDEJS_KERNEL_GAUSSIAN)
                k_density_dense32<1><<<grid, kD32Threads, 0, st>>>(wd, g, sl.scaled.as<double2>(),
                                                                  sl.dense_partial.as<double>());
            else
                k_density_dense32<0><<<grid, kD32Threads, 0, st>>>(wd, g, sl.scaled.as<double2>(),
                                                                  sl.dense_partial.as<double>());
        } else {
            const int ntd = (p.R + kDenseTile - 1) / kDenseTile;
            nslices = std::max(1, std::min(64, (4 * c.sm_count) / std::max(1, ntd * ntd * n_sets)));
            OKR(sl.dense_partial.ensure((size_t)nslices * n_sets * G * sizeof(double)));
            KTimer t(c, 4, st);
            dim3 grid(ntd * ntd, n_sets, nslices);
            if (p.kernel == KDEJS_KERNEL_GAUSSIAN)
                k_density_dense64<1><<<grid, 256, 0, st>>>(wd, g, sl.scaled.as<double2>(),
                                                           sl.dense_partial.as<double>(), nslices);
            else
                k_density_dense64<0><<<grid, 256, 0, st>>>(wd, g, sl.scaled.as<double2>(),
                                                           sl.dense_partial.as<double>(), nslices);
        }
        {
            KTimer t(c, 4, st);
            k_dense_finish<<<dim3(nparts, n_sets), 256, 0, st>>>(
                wd, g, sl.dense_partial.as<double>(), nslices, sl.dens.as<double>(),
                sl.pz.as<double>(), sl.tilesum.as<double>(), nparts);
        }
    }
    if (!io.density_only && !io.shard) {
        OKR(sl.totals.ensure(sizeof(double) * kMaxWaveSets));
        {
            KTimer t(c, 5, st);
            k_set_totals<<<n_sets, kJsThreads, 0, st>>>(nparts, sl.tilesum.as<double>(), sl.totals.as<double>());
        }
        KTimer t(c, 5, st);
        if (lazy_empty)
            k_js_tiles<<<dim3(g.NT, n_items), kJsThreads, 0, st>>>(
                wd, g, sl.pz.as<double>(), sl.totals.as<double>(), sl.rstat.as<RawStat>(),
                sl.tileflag.as<uint8_t>(), sl.jspart.as<double>(), tk_js, io.out_js, io.out_status, io.item_base);
        else
        k_js<<<dim3(js_blocks, n_items), kJsThreads, 0, st>>>(
            wd, (int)G, nparts, 0, (int)G, sl.pz.as<double>(), sl.totals.as<double>(),
            sl.rstat.as<RawStat>(), io.want_z ? sl.zout.as<double>() : nullptr,
            sl.jspart.as<double>(), tk_js, io.out_js, io.out_status, nullptr, nullptr, io.item_base);
    }
    CU(cudaGetLastError());
    return KDEJS_OK;
}

int ensure_host_results(Ctx &c, size_t n) {
    if (n <= c.h_cap) return KDEJS_OK;
    if (c.h_js) { cudaFreeHost(c.h_js); cudaFreeHost(c.h_status); c.h_js = nullptr; c.h_status = nullptr; }
    size_t want = n + n / 2 + 64;
    CU(cudaMallocHost(&c.h_js, want * sizeof(double)));
    CU(cudaMallocHost(&c.h_status, want * sizeof(int32_t)));
    c.h_cap = want;
    return KDEJS_OK;
}

int wave_capacity(const Params &p, int64_t pts_per_item) {
    // keep a wave's scratch under ~3 GB; at the headline size (200k points, R=256) this is 32
    double per_item = (double)pts_per_item * (16 + 16 + 2) + 2.0 * p.R * p.R * 8 * 3 + 2.0 * 150 * 4096 * 4;
    if (p.algo == KDEJS_ALGO_DENSE)     // fp64 slice partials: one grid per 1024 points per set
        per_item += 2.0 * ((double)pts_per_item / 2 / kD32Slice + 1) * p.R * p.R * 8;
    int cap = (int)std::max(1.0, std::floor(3.0e9 / per_item));
    return std::min(cap, kMaxWaveItems);
}

// Pairs over a pool of raw DEVICE sets; items [0, n_items): (ia[i], ib[i]).  Runs on `st`.
int run_pairs_dev(Ctx &c, cudaStream_t st, const Params &p, const std::vector<RawSet> &pool,
                  const int32_t *ia, const int32_t *ib, int64_t n_items, double *out_js_dev,
                  int32_t *status_dev, int64_t out_base, bool want_z) {
    GridGeom g = make_geom(p);
    // Two choices of run_wave() depend on the sizes of the sets it is given: moment cells (fine grids: at least 128
    // points per cell on average) and the fp64 kernel for sets too small for the fp32 error bound (< 64 points).  An
    // evaluation must get the same arithmetic whatever else travels in its wave -- batch results equal single calls
    // bit for bit -- so both are decided per EVALUATION here and a wave only holds evaluations of one class.
    const int64_t ncells_g = (int64_t)g.C * g.C;
    auto eval_class = [&](int64_t k) {
        const int64_t na = pool[ia[k]].n, nb = pool[ib[k]].n;
        int cls = 0;
        if (std::min(na, nb) < kMinPoints32) cls |= 1;
        if (g.use_moments == 1 && na + nb >= (int64_t)128 * 2 * ncells_g) cls |= 2;
        return cls;
    };
    int64_t i = 0;
    while (i < n_items) {
        WaveDesc wd;
        memset(&wd, 0, sizeof wd);
        int local_of_global_n = 0;
        std::vector<std::pair<int, int>> seen;   // (global raw, local raw)
        auto local = [&](int gidx) {
            for (auto &pr : seen) if (pr.first == gidx) return pr.second;
            wd.raw[local_of_global_n] = pool[gidx];
            seen.push_back({gidx, local_of_global_n});
            return local_of_global_n++;
        };
        int64_t maxpts = 1;
        for (int64_t k = i; k < std::min(n_items, i + kMaxWaveItems); ++k)
            maxpts = std::max<int64_t>(maxpts, pool[ia[k]].n + pool[ib[k]].n);
        const int cap = wave_capacity(p, maxpts);
        int items = 0;
        const int wave_class = eval_class(i);
        while (items < cap && i + items < n_items && eval_class(i + items) == wave_class) {
            const int a = ia[i + items], b = ib[i + items];
            const int la = local(a), lb = local(b);
            SetDesc &s0 = wd.set[2 * items], &s1 = wd.set[2 * items + 1];
            s0.rs_self = la; s0.rs_other = lb; s0.n = (int32_t)pool[a].n;
            s1.rs_self = lb; s1.rs_other = la; s1.n = (int32_t)pool[b].n;
            ++items;
        }
        wd.n_raw = local_of_global_n;
        wd.n_sets = 2 * items;
        WaveIO io;
        io.out_js = out_js_dev; io.out_status = status_dev; io.item_base = (int)(out_base + i); io.want_z = want_z;
        io.force_moments = (wave_class & 2) ? 1 : 0;          // (ignored unless the geometry is merely eligible)
        OKR(run_wave(c, c.S(), st, p, g, wd, io));
        i += items;
    }
    return KDEJS_OK;
}

int check_set(const double *p, int64_t n, const char *what) {
    if (!p) return fail(KDEJS_ERR_ARG, std::string(what) + ": null pointer");
    if (n <= 0) return fail(KDEJS_ERR_ARG, std::string("Found array with 0 sample(s) (") + what + ") while a minimum of 1 is required.");
    if (n > 0x7fffffffLL / 4) return fail(KDEJS_ERR_ARG, std::string(what) + ": too many points for one set");
    return KDEJS_OK;
}

int finish_status(const int32_t *status, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (status[i] == KDEJS_ERR_INF)
            return fail(KDEJS_ERR_INF, "Input X contains infinity or a value too large for dtype('float64').");
    return KDEJS_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int kdejs_version(void) { return KDEJS_VERSION; }

const char *kdejs_last_error(void) { return g_err.c_str(); }

int kdejs_device_count(int *count) {
    if (!count) return fail(KDEJS_ERR_ARG, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(KDEJS_ERR_CUDA, cudaGetErrorString(e)); }
    *count = n;
    return KDEJS_OK;
}

int kdejs_init(int device) {
    Ctx *c;
    return get_ctx(device, &c);
}

int kdejs_device_numa_node(int device, int *node) {
    if (!node) return fail(KDEJS_ERR_ARG, "node is null");
    *node = -1;
    char id[64] = {0};
    cudaError_t e = cudaDeviceGetPCIBusId(id, (int)sizeof id - 1, device);
    if (e != cudaSuccess) return fail(KDEJS_ERR_CUDA, cudaGetErrorString(e));
    for (char *q = id; *q; ++q) *q = (char)tolower((unsigned char)*q);     // sysfs spells bus ids in lower case
    const std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
    if (FILE *f = fopen(path.c_str(), "r")) {
        int v = -1;
        if (fscanf(f, "%d", &v) == 1) *node = v;
        fclose(f);
    }
    return KDEJS_OK;
}

void kdejs_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    for (auto &c : g_ctx) {
        if (!c.ready.load(std::memory_order_acquire)) continue;
        std::lock_guard<std::mutex> lc(c.mu);          // no entry point is inside this context while it is torn down
        cudaSetDevice(c.dev);
        cudaDeviceSynchronize();
        destroy_ctx(c);
        c.ready.store(false, std::memory_order_release);
    }
}

void *kdejs_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void kdejs_host_free(void *p) { if (p) cudaFreeHost(p); }

int kdejs_discrepancy_pairs(int device, const double *const *sets, const int64_t *n_pts, int n_sets,
                            const int32_t *ia, const int32_t *ib, int64_t n_pairs, int resolution,
                            double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                            double *out_js) {
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(check_params(p));
    if (!sets || !n_pts || n_sets < 1) return fail(KDEJS_ERR_ARG, "sets/n_pts missing");
    if (n_pairs < 0 || (n_pairs > 0 && (!ia || !ib || !out_js))) return fail(KDEJS_ERR_ARG, "pair arrays missing");
    for (int s = 0; s < n_sets; ++s) OKR(check_set(sets[s], n_pts[s], "point set"));
    for (int64_t k = 0; k < n_pairs; ++k)
        if (ia[k] < 0 || ia[k] >= n_sets || ib[k] < 0 || ib[k] >= n_sets) return fail(KDEJS_ERR_ARG, "pair index out of range");
    if (n_pairs == 0) return KDEJS_OK;
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(acquire_scratch(*c, c->s_compute));
    // upload the pool once
    std::vector<RawSet> pool(n_sets);
    size_t total = 0;
    std::vector<size_t> offs(n_sets);
    for (int s = 0; s < n_sets; ++s) { offs[s] = total; total += ((size_t)n_pts[s] * 16 + 255) / 256 * 256; }
    OKR(c->pool_raw.ensure(total));
    std::vector<Stager::Range> up(n_sets);
    for (int s = 0; s < n_sets; ++s) {
        up[s] = Stager::Range{(char *)c->pool_raw.p + offs[s], sets[s], (size_t)n_pts[s] * 16};
        pool[s] = RawSet{(const double *)((char *)c->pool_raw.p + offs[s]), n_pts[s]};
    }
    OKR(h2d_ranges(device, up, c->s_compute));
    OKR(c->js_out.ensure(sizeof(double) * (size_t)n_pairs));
    OKR(c->status_out.ensure(sizeof(int32_t) * (size_t)n_pairs));
    OKR(ensure_host_results(*c, (size_t)n_pairs));
    OKR(run_pairs_dev(*c, c->s_compute, p, pool, ia, ib, n_pairs, c->js_out.as<double>(),
                      c->status_out.as<int32_t>(), 0, false));
    CU(cudaMemcpyAsync(c->h_js, c->js_out.p, sizeof(double) * (size_t)n_pairs, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaMemcpyAsync(c->h_status, c->status_out.p, sizeof(int32_t) * (size_t)n_pairs, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    release_scratch(*c, c->s_compute, true);
    memcpy(out_js, c->h_js, sizeof(double) * (size_t)n_pairs);
    return finish_status(c->h_status, n_pairs);
}

int kdejs_discrepancy(int device, const double *a, int64_t na, const double *b, int64_t nb, int resolution,
                      double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                      double *out_js, double *out_z1, double *out_z2, double *out_d1, double *out_d2) {
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(check_params(p));
    OKR(check_set(a, na, "df1"));
    OKR(check_set(b, nb, "df2"));
    if (!out_js) return fail(KDEJS_ERR_ARG, "out_js is null");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t offb = ((size_t)na * 16 + 255) / 256 * 256;
    OKR(c->pool_raw.ensure(offb + (size_t)nb * 16));
    cudaStream_t st = c->s_compute;
    OKR(acquire_scratch(*c, st));
    OKR(h2d_ranges(device, {Stager::Range{c->pool_raw.p, a, (size_t)na * 16},
                            Stager::Range{(char *)c->pool_raw.p + offb, b, (size_t)nb * 16}}, st));
    std::vector<RawSet> pool = {RawSet{(const double *)c->pool_raw.p, na},
                                RawSet{(const double *)((char *)c->pool_raw.p + offb), nb}};
    OKR(c->js_out.ensure(sizeof(double)));
    OKR(c->status_out.ensure(sizeof(int32_t)));
    OKR(ensure_host_results(*c, 1));
    const int32_t ia = 0, ib = 1;
    const bool want_z = out_z1 || out_z2 || out_d1 || out_d2;     // any grid output: every node value is materialised
    OKR(run_pairs_dev(*c, st, p, pool, &ia, &ib, 1, c->js_out.as<double>(), c->status_out.as<int32_t>(), 0, want_z));
    CU(cudaMemcpyAsync(c->h_js, c->js_out.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_status, c->status_out.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    const size_t G = (size_t)resolution * resolution;
    Slot &sl = c->S();
    if (out_z1) CU(cudaMemcpyAsync(out_z1, sl.zout.as<double>(), G * 8, cudaMemcpyDeviceToHost, st));
    if (out_z2) CU(cudaMemcpyAsync(out_z2, sl.zout.as<double>() + G, G * 8, cudaMemcpyDeviceToHost, st));
    if (out_d1) CU(cudaMemcpyAsync(out_d1, sl.dens.as<double>(), G * 8, cudaMemcpyDeviceToHost, st));
    if (out_d2) CU(cudaMemcpyAsync(out_d2, sl.dens.as<double>() + G, G * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    release_scratch(*c, st, true);
    *out_js = c->h_js[0];
    return finish_status(c->h_status, 1);
}

namespace {
struct SummaryArgs {          // kdejs_summary_batch: what is fused behind the discrepancies of the batch
    const double *const *freqs; const int64_t *n_freqs; double core_thr, rare_thr; const double *observed;
    double *out_summary, *out_d, *out_log_d;
};
}  // namespace

static int batch_impl(int device, const double *obs, int64_t n_obs, const double *const *sims,
                      const int64_t *n_sims, int B, int resolution, double bandwidth, double gamma,
                      double eps, int log_flag, int kernel, int algo, double *out_js, const SummaryArgs *sum) {
    NvtxRange nvtx_call(sum ? "kdejs_summary_batch" : "kdejs_discrepancy_batch");
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(check_params(p));
    OKR(check_set(obs, n_obs, "obs"));
    if (B < 0 || (B > 0 && (!sims || !n_sims))) return fail(KDEJS_ERR_ARG, "batch arrays missing");
    for (int i = 0; i < B; ++i) OKR(check_set(sims[i], n_sims[i], "sim"));
    if (B == 0) return KDEJS_OK;
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(acquire_scratch(*c, c->s_compute));
    OKR(c->pool_raw.ensure((size_t)n_obs * 16));
    OKR(c->js_out.ensure(sizeof(double) * (size_t)B));
    OKR(c->status_out.ensure(sizeof(int32_t) * (size_t)B));
    OKR(ensure_host_results(*c, (size_t)B));
    int64_t nmax = 1;
    for (int i = 0; i < B; ++i) nmax = std::max(nmax, n_sims[i]);
    // H2D and kernels run on two streams.  The raw sims land in a ring of device buffers (up to 2 GB) that
    // the copy stream fills as far ahead as it can; the kernels of a wave wait only for that wave's copies.
    // The first wave is short (its copy is the only one that cannot hide behind kernels), later ones are
    // large enough to run the kernels efficiently.  All waves share one scratch slot: they are ordered on
    // the compute stream anyway.
    const int cap = wave_capacity(p, n_obs + nmax);
    const size_t sim_stride = ((size_t)nmax * 16 + 255) / 256 * 256;
    const int ring_items = (int)std::min<int64_t>(B, std::max<int64_t>(cap, (int64_t)(2.0e9 / (double)sim_stride)));
    OKR(c->sim_ring.ensure(sim_stride * (size_t)ring_items));
    // Pageable sources (ordinary numpy arrays) are staged through pinned slots by worker threads that run ahead
    // of this loop (kdejs_stager.h); page-locked sources are copied directly.
    std::vector<char> staged(B + 1, 0);
    Stager *sg = nullptr;
    {
        std::vector<Stager::Range> plan;
        size_t pageable_bytes = 0;
        if (Stager::pageable(obs)) { staged[B] = 1; pageable_bytes += (size_t)n_obs * 16; plan.push_back(Stager::Range{c->pool_raw.p, obs, (size_t)n_obs * 16}); }
        for (int i = 0; i < B; ++i) {
            // a sim that starts where its page-locked predecessor ends (one arena) is not asked about again: a wrong
            // guess would only cost speed (cudaMemcpyAsync accepts pageable sources), and 64 queries are ~30 us of
            // the critical path before the first copy
            if (i > 0 && !staged[i - 1] && (const char *)sims[i] == (const char *)sims[i - 1] + (size_t)n_sims[i - 1] * 16) continue;
            if (Stager::pageable(sims[i])) {
                staged[i] = 1;
                pageable_bytes += (size_t)n_sims[i] * 16;
                plan.push_back(Stager::Range{(char *)c->sim_ring.p + sim_stride * (size_t)(i % ring_items), sims[i], (size_t)n_sims[i] * 16});
            }
        }
        if (pageable_bytes >= kStageMinBytes) sg = Stager::get(device);
        if (sg) sg->plan(plan);
    }
    if (sg && staged[B]) CU(sg->issue_range(c->s_compute));
    else CU(cudaMemcpyAsync(c->pool_raw.p, obs, (size_t)n_obs * 16, cudaMemcpyHostToDevice, c->s_compute));
    int wave_size = std::min(cap, 16);
    if (const char *e = getenv("KDEJS_HOST_WAVE")) wave_size = std::max(1, std::min(cap, atoi(e)));
    int first_wave = 4, tail_wave = 0;
    if (const char *e = getenv("KDEJS_FIRST_WAVE")) first_wave = std::max(1, atoi(e));
    if (const char *e = getenv("KDEJS_TAIL_WAVE")) tail_wave = std::max(0, std::min(cap, atoi(e)));
    std::vector<int> wave_plan;                       // experiments: explicit wave sizes, the last one repeats
    if (const char *e = getenv("KDEJS_WAVE_PLAN"))
        for (const char *q = e; *q;) { wave_plan.push_back(std::max(1, std::min(cap, atoi(q)))); while (*q && *q != ',') ++q; if (*q) ++q; }
    static const bool merge_copies = !(getenv("KDEJS_MERGE_COPIES") && atoi(getenv("KDEJS_MERGE_COPIES")) == 0);
    std::vector<int> wave_of_item(B);
    std::vector<cudaEvent_t> ev_copied, ev_done;
    auto new_event = [&](cudaEvent_t &e) -> int {
        if (!c->free_plain_events.empty()) { e = c->free_plain_events.back(); c->free_plain_events.pop_back(); return KDEJS_OK; }
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return KDEJS_OK;
    };
    int rc = KDEJS_OK;
    int wave = 0;
    for (int i0 = 0; i0 < B && rc == KDEJS_OK; ++wave) {
        int items = std::min(wave == 0 ? std::min(cap, first_wave) : wave_size, B - i0);
        // the kernels of the last wave run after the last copy has landed: keep that wave short (tail_wave items)
        if (tail_wave > 0 && B - i0 > tail_wave && B - i0 - items < tail_wave) items = B - i0 - tail_wave;
        if (!wave_plan.empty()) items = std::min(wave_plan[std::min<size_t>(wave, wave_plan.size() - 1)], B - i0);
        cudaEvent_t ec, ed;
        if ((rc = new_event(ec)) != KDEJS_OK || (rc = new_event(ed)) != KDEJS_OK) break;
        ev_copied.push_back(ec);
        ev_done.push_back(ed);
        std::vector<RawSet> pool(items + 1);
        pool[0] = RawSet{(const double *)c->pool_raw.p, n_obs};
        std::vector<int32_t> ia(items, 0), ib(items);
        for (int k = 0; k < items; ++k) {
            const int i = i0 + k;
            wave_of_item[i] = wave;
            if (i >= ring_items)        // ring slot reuse: the evaluation that read it must have finished
                if (cudaEventSynchronize(ev_done[wave_of_item[i - ring_items]]) != cudaSuccess) { rc = fail(KDEJS_ERR_CUDA, "event wait failed"); break; }
            char *dst = (char *)c->sim_ring.p + sim_stride * (size_t)(i % ring_items);
            pool[k + 1] = RawSet{(const double *)dst, n_sims[i]};
            ib[k] = k + 1;
        }
        if (rc != KDEJS_OK) break;
        // One cudaMemcpyAsync per run of sims that are contiguous on the host AND land contiguously in the ring
        // (a caller that carved its sims out of one pinned arena, PinnedPool.empty_batch): a 1.6 MB copy reaches
        // ~2/3 of the link rate, a 26 MB one all of it.  Anything else: one copy per sim, as before.
        for (int k = 0; k < items && rc == KDEJS_OK;) {
            const int i = i0 + k;
            char *dst = (char *)c->sim_ring.p + sim_stride * (size_t)(i % ring_items);
            size_t bytes = (size_t)n_sims[i] * 16;
            int m = 1;
            if (!(sg && staged[i]) && merge_copies)
                while (k + m < items && !(sg && staged[i + m]) && bytes == sim_stride * (size_t)m &&
                       (const char *)sims[i + m] == (const char *)sims[i] + bytes &&
                       (i + m) % ring_items == (i % ring_items) + m) {
                    bytes += (size_t)n_sims[i + m] * 16;
                    ++m;
                }
            const cudaError_t ce = (sg && staged[i]) ? sg->issue_range(c->s_copy)
                                                     : cudaMemcpyAsync(dst, sims[i], bytes, cudaMemcpyHostToDevice, c->s_copy);
            if (ce != cudaSuccess) rc = fail(KDEJS_ERR_CUDA, "H2D copy of a simulated set failed");
            k += m;
        }
        if (rc != KDEJS_OK) break;
        cudaEventRecord(ec, c->s_copy);
        cudaStreamWaitEvent(c->s_compute, ec, 0);
        rc = run_pairs_dev(*c, c->s_compute, p, pool, ia.data(), ib.data(), items,
                           c->js_out.as<double>(), c->status_out.as<int32_t>(), i0, false);
        cudaEventRecord(ed, c->s_compute);
        i0 += items;
    }
    std::vector<double> h_freqs, h_sum;
    std::vector<int64_t> h_offs;
    if (rc == KDEJS_OK && sum) {
        // gene frequencies of all simulations as one upload, the summary kernel behind the last wave, one read-back
        h_offs.assign(B + 1, 0);
        for (int i = 0; i < B; ++i) h_offs[i + 1] = h_offs[i] + sum->n_freqs[i];
        h_freqs.resize((size_t)std::max<int64_t>(h_offs[B], 1));
        for (int i = 0; i < B; ++i) memcpy(h_freqs.data() + h_offs[i], sum->freqs[i], sizeof(double) * (size_t)sum->n_freqs[i]);
        const size_t off_bytes = ((size_t)(B + 1) * 8 + 255) / 256 * 256, fr_bytes = (h_freqs.size() * 8 + 255) / 256 * 256;
        if ((rc = c->summary.ensure(off_bytes + fr_bytes + (size_t)B * 48)) == KDEJS_OK) {
            char *base = c->summary.as<char>();
            h_sum.resize((size_t)B * 6);
            if (cudaMemcpyAsync(base, h_offs.data(), (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess ||
                cudaMemcpyAsync(base + off_bytes, h_freqs.data(), h_freqs.size() * 8, cudaMemcpyHostToDevice, c->s_compute) != cudaSuccess)
                rc = fail(KDEJS_ERR_CUDA, "H2D copy of the gene frequencies failed");
            else {
                {
                    KTimer t(*c, 9, c->s_compute);
                    k_summary<<<B, 256, 0, c->s_compute>>>(
                        reinterpret_cast<const double *>(base + off_bytes), reinterpret_cast<const int64_t *>(base),
                        sum->core_thr, sum->rare_thr, c->js_out.as<double>(), sum->observed[0], sum->observed[1],
                        sum->observed[2], sum->observed[3], reinterpret_cast<double *>(base + off_bytes + fr_bytes));
                }
                if (cudaMemcpyAsync(h_sum.data(), base + off_bytes + fr_bytes, (size_t)B * 48, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess)
                    rc = fail(KDEJS_ERR_CUDA, "D2H copy of the summary vectors failed");
            }
        }
    }
    if (rc == KDEJS_OK) {
        if (cudaMemcpyAsync(c->h_js, c->js_out.p, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess ||
            cudaMemcpyAsync(c->h_status, c->status_out.p, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess)
            rc = fail(KDEJS_ERR_CUDA, "D2H copy of the results failed");
    }
    if (sg && rc != KDEJS_OK) sg->abandon();
    cudaError_t se = cudaStreamSynchronize(c->s_compute);
    cudaStreamSynchronize(c->s_copy);
    release_scratch(*c, c->s_compute, true);
    for (auto e : ev_copied) c->free_plain_events.push_back(e);
    for (auto e : ev_done) c->free_plain_events.push_back(e);
    if (rc != KDEJS_OK) return rc;
    if (se != cudaSuccess) return fail(KDEJS_ERR_CUDA, std::string("kernel execution failed: ") + cudaGetErrorString(se));
    if (out_js) memcpy(out_js, c->h_js, sizeof(double) * (size_t)B);
    if (sum)
        for (int i = 0; i < B; ++i) {
            memcpy(sum->out_summary + 4 * (size_t)i, h_sum.data() + 6 * (size_t)i, 32);
            if (sum->out_d) sum->out_d[i] = h_sum[6 * (size_t)i + 4];
            if (sum->out_log_d) sum->out_log_d[i] = h_sum[6 * (size_t)i + 5];
        }
    return finish_status(c->h_status, B);
}

int kdejs_tsv_parse_text(int device, const char *text, int64_t bytes, int skip_rows, double *out_pts,
                         int64_t capacity_rows, int64_t *out_rows, int32_t *out_status) {
    if (!text || bytes < 0 || !out_rows || !out_status || (capacity_rows > 0 && !out_pts))
        return fail(KDEJS_ERR_ARG, "null argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = c->s_compute;
    OKR(acquire_scratch(*c, st));
    const size_t text_stride = ((size_t)bytes + 64 + 255) / 256 * 256;
    const int64_t begin = kdejs_tsv_data_offset(text, bytes, skip_rows);
    const int chunks = std::max<int>(1, (int)((bytes - begin / kTsvChunk * kTsvChunk + kTsvChunk - 1) / kTsvChunk));
    OKR(c->tsv_text[0].ensure(text_stride));
    OKR(c->tsv_chunkrows[0].ensure((size_t)chunks * sizeof(uint32_t)));
    OKR(c->tsv_rows[0].ensure(sizeof(int64_t)));
    OKR(c->tsv_status.ensure(sizeof(unsigned)));
    if (!c->h_tsv_rows) CU(cudaMallocHost(&c->h_tsv_rows, 2 * kMaxTsvFiles * sizeof(int64_t)));
    TsvWave wv;
    memset(&wv, 0, sizeof wv);
    wv.f[0] = TsvFile{c->tsv_text[0].as<char>(), begin, bytes, nullptr, 0};
    if (bytes > 0) CU(cudaMemcpyAsync(c->tsv_text[0].p, text, (size_t)bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->tsv_status.p, 0, sizeof(unsigned), st));
    k_tsv_count<<<dim3(chunks, 1), kTsvThreads, 0, st>>>(wv, chunks, c->tsv_chunkrows[0].as<uint32_t>());
    k_tsv_scan<<<1, 1024, 0, st>>>(chunks, c->tsv_chunkrows[0].as<uint32_t>(), c->tsv_rows[0].as<int64_t>());
    CU(cudaMemcpyAsync(c->h_tsv_rows, c->tsv_rows[0].p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int64_t rows = c->h_tsv_rows[0];
    *out_rows = rows;
    *out_status = KDEJS_TABLE_OK;
    c->launches += 2;
    int rc = KDEJS_OK;
    if (rows > capacity_rows) rc = fail(KDEJS_ERR_CAPACITY, std::to_string(rows) + " data rows do not fit the buffer");
    else if (rows > 0) {
        if ((rc = c->tsv_sims.ensure((size_t)rows * 16)) == KDEJS_OK) {
            wv.f[0].out = c->tsv_sims.as<double2>();
            wv.f[0].cap_rows = rows;
            k_tsv_parse<<<dim3(chunks, 1), kTsvThreads, 0, st>>>(wv, chunks, c->tsv_chunkrows[0].as<uint32_t>(),
                                                                 c->tsv_status.as<unsigned>(), nullptr);
            c->launches += 1;
            unsigned flags = 0u;
            if (cudaMemcpyAsync(out_pts, c->tsv_sims.p, (size_t)rows * 16, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaMemcpyAsync(&flags, c->tsv_status.p, sizeof flags, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess)
                rc = fail(KDEJS_ERR_CUDA, "device table parser failed");
            else if (flags & kTsvBad) *out_status = KDEJS_TABLE_HOST;
        }
    } else *out_status = KDEJS_TABLE_HOST;
    release_scratch(*c, st, true);
    return rc;
}

int kdejs_discrepancy_batch_text(int device, const double *obs, int64_t n_obs, const char *const *texts,
                                 const int64_t *text_bytes, int B, int skip_rows, int resolution, double bandwidth,
                                 double gamma, double eps, int log_flag, int kernel, int algo, double *out_js,
                                 int64_t *out_rows, int32_t *out_status, double *out_bounds) {
    NvtxRange nvtx_call("kdejs_discrepancy_batch_text");
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(check_params(p));
    OKR(check_set(obs, n_obs, "obs"));
    if (B < 0 || (B > 0 && (!texts || !text_bytes || !out_js || !out_rows || !out_status)))
        return fail(KDEJS_ERR_ARG, "batch arrays missing");
    int64_t max_bytes = 0;
    for (int i = 0; i < B; ++i) {
        if (!texts[i] || text_bytes[i] < 0) return fail(KDEJS_ERR_ARG, "table text missing");
        max_bytes = std::max(max_bytes, text_bytes[i]);
    }
    if (B == 0) return KDEJS_OK;
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(acquire_scratch(*c, c->s_compute));
    int W = 16;                                             // tables per wave
    if (const char *e = getenv("KDEJS_TEXT_WAVE")) W = std::max(1, std::min(std::min(kMaxTsvFiles, kMaxWaveItems), atoi(e)));
    const size_t text_stride = ((size_t)max_bytes + 64 + 255) / 256 * 256;      // padded: the parser reads whole 32-byte segments
    const int chunks_cap = (int)(text_stride / kTsvChunk) + 2;
    for (int k = 0; k < 2; ++k) {
        OKR(c->tsv_text[k].ensure(text_stride * W));
        OKR(c->tsv_chunkrows[k].ensure((size_t)W * chunks_cap * sizeof(uint32_t)));
        OKR(c->tsv_rows[k].ensure(W * sizeof(int64_t)));
    }
    const size_t keys_off = ((size_t)B * sizeof(unsigned) + 255) / 256 * 256;
    OKR(c->tsv_status.ensure(keys_off + (size_t)B * 4 * sizeof(unsigned long long)));
    unsigned long long *d_keys = reinterpret_cast<unsigned long long *>(c->tsv_status.as<char>() + keys_off);
    if (!c->h_tsv_rows) CU(cudaMallocHost(&c->h_tsv_rows, 2 * kMaxTsvFiles * sizeof(int64_t)));
    OKR(c->pool_raw.ensure((size_t)n_obs * 16));
    OKR(c->js_out.ensure(sizeof(double) * (size_t)B));
    OKR(c->status_out.ensure(sizeof(int32_t) * (size_t)B));
    OKR(ensure_host_results(*c, (size_t)B));
    OKR(h2d_ranges(device, {Stager::Range{c->pool_raw.p, obs, (size_t)n_obs * 16}}, c->s_compute));
    CU(cudaMemsetAsync(c->tsv_status.p, 0, keys_off + (size_t)B * 4 * sizeof(unsigned long long), c->s_compute));
    std::vector<int64_t> begin(B);
    for (int i = 0; i < B; ++i) begin[i] = kdejs_tsv_data_offset(texts[i], text_bytes[i], skip_rows);
    // Waves of W tables; the LAST one is short: its count, the host's look at the row totals, its parse and its
    // evaluation all run after the last byte has crossed the link, with nothing left to overlap them.
    int tail = std::min(8, W);              // measured, 64 tables per call: 4.40 ms without, 4.33 with 4, 4.26 with 8
    if (const char *e = getenv("KDEJS_TEXT_TAIL")) tail = std::max(0, std::min(W, atoi(e)));
    std::vector<int> wave_first;                  // first table of every wave, then B
    for (int i = 0; i < B;) {
        wave_first.push_back(i);
        const int rem = B - i;
        i += (tail > 0 && rem > tail && rem <= W) ? rem - tail : std::min(W, rem);
    }
    wave_first.push_back(B);
    const int n_waves = (int)wave_first.size() - 1;
    std::vector<cudaEvent_t> ev_counted(n_waves), ev_parsed(n_waves);
    auto new_event = [&](cudaEvent_t &e) -> int {
        if (!c->free_plain_events.empty()) { e = c->free_plain_events.back(); c->free_plain_events.pop_back(); return KDEJS_OK; }
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return KDEJS_OK;
    };
    for (int k = 0; k < n_waves; ++k) { OKR(new_event(ev_counted[k])); OKR(new_event(ev_parsed[k])); }
    std::vector<TsvWave> waves(n_waves);
    std::vector<int> wave_chunks(n_waves, 1);
    // copy stream: the text of wave k, its line count per 8 KB chunk, the scan, the row totals back to the host
    auto enqueue_copy_count = [&](int k) -> int {
        const int i0 = wave_first[k], nf = wave_first[k + 1] - i0, buf = k & 1;
        if (k >= 2) CU(cudaStreamWaitEvent(c->s_copy, ev_parsed[k - 2], 0));       // the parser is done with this buffer
        TsvWave &wv = waves[k];
        memset(&wv, 0, sizeof wv);
        int chunks = 1;
        for (int j = 0; j < nf; ++j) {
            const int i = i0 + j;
            char *dst = c->tsv_text[buf].as<char>() + text_stride * (size_t)j;
            if (text_bytes[i] > 0) CU(cudaMemcpyAsync(dst, texts[i], (size_t)text_bytes[i], cudaMemcpyHostToDevice, c->s_copy));
            wv.f[j] = TsvFile{dst, begin[i], text_bytes[i], nullptr, 0};
            chunks = std::max<int>(chunks, (int)((text_bytes[i] - begin[i] / kTsvChunk * kTsvChunk + kTsvChunk - 1) / kTsvChunk));
        }
        wave_chunks[k] = chunks;
        {
            KTimer t(*c, 10, c->s_copy);
            k_tsv_count<<<dim3(chunks, nf), kTsvThreads, 0, c->s_copy>>>(wv, chunks, c->tsv_chunkrows[buf].as<uint32_t>());
            k_tsv_scan<<<nf, 1024, 0, c->s_copy>>>(chunks, c->tsv_chunkrows[buf].as<uint32_t>(), c->tsv_rows[buf].as<int64_t>());
        }
        c->launches += 1;
        CU(cudaMemcpyAsync(c->h_tsv_rows + (size_t)buf * kMaxTsvFiles, c->tsv_rows[buf].p, (size_t)nf * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, c->s_copy));
        CU(cudaEventRecord(ev_counted[k], c->s_copy));
        CU(cudaGetLastError());
        return KDEJS_OK;
    };
    int rc = enqueue_copy_count(0);
    std::vector<int> item_of_slot;            // results are written densely: slot -> table
    item_of_slot.reserve(B);
    for (int k = 0; k < n_waves && rc == KDEJS_OK; ++k) {
        const int i0 = wave_first[k], nf = wave_first[k + 1] - i0, buf = k & 1;
        if (k + 1 < n_waves && (rc = enqueue_copy_count(k + 1)) != KDEJS_OK) break;
        if (cudaEventSynchronize(ev_counted[k]) != cudaSuccess) { rc = fail(KDEJS_ERR_CUDA, "table upload / count failed"); break; }
        const int64_t *rows = c->h_tsv_rows + (size_t)buf * kMaxTsvFiles;
        size_t total = 0;
        std::vector<size_t> off(nf);
        for (int j = 0; j < nf; ++j) {
            out_rows[i0 + j] = rows[j];
            off[j] = total;
            total += ((size_t)std::max<int64_t>(rows[j], 0) * 16 + 255) / 256 * 256;
        }
        if ((rc = c->tsv_sims.ensure(std::max<size_t>(total, 256))) != KDEJS_OK) break;
        TsvWave &wv = waves[k];
        std::vector<RawSet> pool(1, RawSet{(const double *)c->pool_raw.p, n_obs});
        std::vector<int32_t> ia, ib;
        const size_t slot0 = item_of_slot.size();
        for (int j = 0; j < nf; ++j) {
            wv.f[j].out = reinterpret_cast<double2 *>(c->tsv_sims.as<char>() + off[j]);
            wv.f[j].cap_rows = rows[j];
            if (rows[j] > 0 && rows[j] <= 0x7fffffffLL / 4) {
                pool.push_back(RawSet{(const double *)wv.f[j].out, rows[j]});
                ia.push_back(0);
                ib.push_back((int32_t)pool.size() - 1);
                item_of_slot.push_back(i0 + j);
            }
        }
        {
            KTimer t(*c, 11, c->s_compute);
            k_tsv_parse<<<dim3(wave_chunks[k], nf), kTsvThreads, 0, c->s_compute>>>(
                wv, wave_chunks[k], c->tsv_chunkrows[buf].as<uint32_t>(), c->tsv_status.as<unsigned>() + i0,
                out_bounds ? d_keys + (size_t)i0 * 4 : nullptr);
        }
        cudaEventRecord(ev_parsed[k], c->s_compute);
        if (!ia.empty())
            rc = run_pairs_dev(*c, c->s_compute, p, pool, ia.data(), ib.data(), (int64_t)ia.size(), c->js_out.as<double>(),
                               c->status_out.as<int32_t>(), (int64_t)slot0, false);
    }
    std::vector<unsigned> tsv_status(B, 0u);
    const size_t n_slots = item_of_slot.size();
    if (rc == KDEJS_OK) {
        if ((n_slots && (cudaMemcpyAsync(c->h_js, c->js_out.p, sizeof(double) * n_slots, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess ||
                         cudaMemcpyAsync(c->h_status, c->status_out.p, sizeof(int32_t) * n_slots, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess)) ||
            cudaMemcpyAsync(tsv_status.data(), c->tsv_status.p, sizeof(unsigned) * (size_t)B, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess ||
            (out_bounds && cudaMemcpyAsync(out_bounds, d_keys, sizeof(double) * 4 * (size_t)B, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess))
            rc = fail(KDEJS_ERR_CUDA, "D2H copy of the results failed");
    }
    cudaError_t se = cudaStreamSynchronize(c->s_compute);
    cudaStreamSynchronize(c->s_copy);
    release_scratch(*c, c->s_compute, true);
    for (auto e : ev_counted) c->free_plain_events.push_back(e);
    for (auto e : ev_parsed) c->free_plain_events.push_back(e);
    if (rc != KDEJS_OK) return rc;
    if (se != cudaSuccess) return fail(KDEJS_ERR_CUDA, std::string("kernel execution failed: ") + cudaGetErrorString(se));
    const double nan = std::nan("");
    for (int i = 0; i < B; ++i) { out_js[i] = nan; out_status[i] = KDEJS_TABLE_HOST; }      // until shown otherwise
    if (out_bounds)          // keys -> doubles: (min, max) per column as np.min / np.max give them (NaN if the column holds one)
        for (int i = 0; i < B; ++i)
            for (int q = 0; q < 4; ++q) {
                unsigned long long k;
                memcpy(&k, out_bounds + 4 * (size_t)i + q, sizeof k);
                if (!(q & 1)) k = ~k;                        // minima were folded as the maximum of ~key
                unsigned long long bits = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
                if (tsv_status[i] & ((q < 2) ? kTsvNan0 : kTsvNan1)) bits = 0x7ff8000000000000ull;
                memcpy(out_bounds + 4 * (size_t)i + q, &bits, sizeof bits);
            }
    for (size_t s = 0; s < n_slots; ++s) {
        const int i = item_of_slot[s];
        if (tsv_status[i] & kTsvBad) continue;              // a field the device parser does not cover: host reader
        if (c->h_status[s] == KDEJS_ERR_INF) { out_status[i] = KDEJS_TABLE_INF; continue; }
        out_js[i] = c->h_js[s];
        out_status[i] = KDEJS_TABLE_OK;
    }
    return KDEJS_OK;
}

int kdejs_discrepancy_batch(int device, const double *obs, int64_t n_obs, const double *const *sims,
                            const int64_t *n_sims, int B, int resolution, double bandwidth, double gamma,
                            double eps, int log_flag, int kernel, int algo, double *out_js) {
    if (B > 0 && !out_js) return fail(KDEJS_ERR_ARG, "batch arrays missing");
    return batch_impl(device, obs, n_obs, sims, n_sims, B, resolution, bandwidth, gamma, eps, log_flag, kernel, algo,
                      out_js, nullptr);
}

int kdejs_summary_batch(int device, const double *obs, int64_t n_obs, const double *const *sims,
                        const int64_t *n_sims, const double *const *freqs, const int64_t *n_freqs, int B,
                        double core_threshold, double rare_threshold, const double observed[4], int resolution,
                        double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                        double *out_summary, double *out_d, double *out_log_d) {
    if (B < 0) return fail(KDEJS_ERR_ARG, "negative batch size");
    if (B == 0) return KDEJS_OK;
    if (!freqs || !n_freqs || !observed || !out_summary) return fail(KDEJS_ERR_ARG, "summary arrays missing");
    for (int i = 0; i < B; ++i)
        if (n_freqs[i] < 0 || (n_freqs[i] > 0 && !freqs[i])) return fail(KDEJS_ERR_ARG, "bad gene-frequency vector");
    SummaryArgs sa{freqs, n_freqs, core_threshold, rare_threshold, observed, out_summary, out_d, out_log_d};
    return batch_impl(device, obs, n_obs, sims, n_sims, B, resolution, bandwidth, gamma, eps, log_flag, kernel, algo,
                      nullptr, &sa);
}

int kdejs_summary_from_js(int device, const double *js, const double *const *freqs, const int64_t *n_freqs, int B,
                          double core_threshold, double rare_threshold, const double observed[4],
                          double *out_summary, double *out_d, double *out_log_d) {
    NvtxRange nvtx_call("kdejs_summary_from_js");
    if (B < 0) return fail(KDEJS_ERR_ARG, "negative batch size");
    if (B == 0) return KDEJS_OK;
    if (!js || !freqs || !n_freqs || !observed || !out_summary) return fail(KDEJS_ERR_ARG, "summary arrays missing");
    for (int i = 0; i < B; ++i)
        if (n_freqs[i] < 0 || (n_freqs[i] > 0 && !freqs[i])) return fail(KDEJS_ERR_ARG, "bad gene-frequency vector");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = c->s_compute;
    OKR(acquire_scratch(*c, st));
    // the layout of batch_impl's summary step: offsets | frequencies | results, and the discrepancies in js_out
    std::vector<int64_t> h_offs((size_t)B + 1, 0);
    for (int i = 0; i < B; ++i) h_offs[i + 1] = h_offs[i] + n_freqs[i];
    std::vector<double> h_freqs((size_t)std::max<int64_t>(h_offs[B], 1));
    for (int i = 0; i < B; ++i)
        if (n_freqs[i] > 0) memcpy(h_freqs.data() + h_offs[i], freqs[i], sizeof(double) * (size_t)n_freqs[i]);
    const size_t off_bytes = ((size_t)(B + 1) * 8 + 255) / 256 * 256, fr_bytes = (h_freqs.size() * 8 + 255) / 256 * 256;
    OKR(c->summary.ensure(off_bytes + fr_bytes + (size_t)B * 48));
    OKR(c->js_out.ensure(sizeof(double) * (size_t)B));
    char *base = c->summary.as<char>();
    std::vector<double> h_sum((size_t)B * 6);
    int rc = KDEJS_OK;
    if (cudaMemcpyAsync(c->js_out.p, js, sizeof(double) * (size_t)B, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(base, h_offs.data(), (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(base + off_bytes, h_freqs.data(), h_freqs.size() * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
        rc = fail(KDEJS_ERR_CUDA, "H2D copy of the discrepancies / gene frequencies failed");
    if (rc == KDEJS_OK) {
        {
            KTimer t(*c, 9, st);
            k_summary<<<B, 256, 0, st>>>(reinterpret_cast<const double *>(base + off_bytes), reinterpret_cast<const int64_t *>(base),
                                        core_threshold, rare_threshold, c->js_out.as<double>(), observed[0], observed[1],
                                        observed[2], observed[3], reinterpret_cast<double *>(base + off_bytes + fr_bytes));
        }
        if (cudaGetLastError() != cudaSuccess ||
            cudaMemcpyAsync(h_sum.data(), base + off_bytes + fr_bytes, (size_t)B * 48, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = fail(KDEJS_ERR_CUDA, "summary kernel / D2H copy of the summary vectors failed");
    }
    const cudaError_t se = cudaStreamSynchronize(st);         // the host vectors above are read until here
    release_scratch(*c, st, true);
    if (rc != KDEJS_OK) return rc;
    if (se != cudaSuccess) return fail(KDEJS_ERR_CUDA, std::string("kernel execution failed: ") + cudaGetErrorString(se));
    for (int i = 0; i < B; ++i) {
        memcpy(out_summary + 4 * (size_t)i, h_sum.data() + 6 * (size_t)i, 32);
        if (out_d) out_d[i] = h_sum[6 * (size_t)i + 4];
        if (out_log_d) out_log_d[i] = h_sum[6 * (size_t)i + 5];
    }
    return KDEJS_OK;
}

int kdejs_discrepancy_batch_dev(int device, const double *obs_dev, int64_t n_obs, const double *const *sims_dev,
                                const int64_t *n_sims, int B, int resolution, double bandwidth, double gamma,
                                double eps, int log_flag, int kernel, int algo, double *out_js_dev,
                                int32_t *status_dev, void *stream) {
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(check_params(p));
    OKR(check_set(obs_dev, n_obs, "obs"));
    if (B < 0 || (B > 0 && (!sims_dev || !n_sims || !out_js_dev))) return fail(KDEJS_ERR_ARG, "batch arrays missing");
    if (((uintptr_t)obs_dev & 15) != 0) return fail(KDEJS_ERR_ARG, "obs_dev must be 16-byte aligned");
    for (int i = 0; i < B; ++i) {
        OKR(check_set(sims_dev[i], n_sims[i], "sim"));
        if (((uintptr_t)sims_dev[i] & 15) != 0) return fail(KDEJS_ERR_ARG, "sims_dev[i] must be 16-byte aligned");
    }
    if (B == 0) return KDEJS_OK;
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    OKR(acquire_scratch(*c, st));
    std::vector<RawSet> pool(B + 1);
    pool[0] = RawSet{obs_dev, n_obs};
    std::vector<int32_t> ia(B, 0), ib(B);
    for (int i = 0; i < B; ++i) { pool[i + 1] = RawSet{sims_dev[i], n_sims[i]}; ib[i] = i + 1; }
    const int rc = run_pairs_dev(*c, st, p, pool, ia.data(), ib.data(), B, out_js_dev, status_dev, 0, false);
    // the call returns with work in flight on the caller's stream: the next user of the scratch -- this stream or
    // another -- is ordered behind it through the event (same-stream users are ordered by the stream itself)
    const int rr = release_scratch(*c, st, false);
    return rc != KDEJS_OK ? rc : rr;
}

int kdejs_kde_density(int device, const double *pts, int64_t n, double gmin, double gmax, int resolution,
                      double bandwidth, int kernel, int algo, double *out_dens) {
    Params p{resolution, bandwidth, 1.0, 0.0, 0, kernel, algo};
    p.gmin = gmin; p.gmax = gmax;
    OKR(check_params(p));
    OKR(check_set(pts, n, "points"));
    if (!out_dens) return fail(KDEJS_ERR_ARG, "out_dens is null");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = c->s_compute;
    OKR(acquire_scratch(*c, st));
    OKR(c->pool_raw.ensure((size_t)n * 16));
    CU(cudaMemcpyAsync(c->pool_raw.p, pts, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    GridGeom g = make_geom(p);
    WaveDesc wd;
    memset(&wd, 0, sizeof wd);
    wd.raw[0] = RawSet{(const double *)c->pool_raw.p, n};
    wd.n_raw = 1; wd.n_sets = 1;
    wd.set[0].rs_self = 0; wd.set[0].rs_other = 0; wd.set[0].n = (int32_t)n;
    WaveIO io;
    io.density_only = true; io.identity_scaler = true;
    OKR(run_wave(*c, c->S(), st, p, g, wd, io));
    CU(cudaMemcpyAsync(out_dens, c->S().dens.p, (size_t)resolution * resolution * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    release_scratch(*c, st, true);
    return KDEJS_OK;
}

static int shard_phase1_core(Ctx *c, const Params &p, RawSet ra, RawSet rb, int row_begin, int row_end,
                             double out_norm[2], int balance_rank = -1, int balance_world = 0,
                             int *out_rows = nullptr) {
    c->shard_valid = false;
    cudaStream_t st = c->s_compute;
    OKR(acquire_scratch(*c, st));
    GridGeom g = make_geom(p);
    g.tile_row_begin = row_begin / kTile;
    g.tile_row_end = (row_end + kTile - 1) / kTile;
    WaveDesc wd;
    memset(&wd, 0, sizeof wd);
    wd.raw[0] = ra;
    wd.raw[1] = rb;
    wd.n_raw = 2; wd.n_sets = 2;
    wd.set[0].rs_self = 0; wd.set[0].rs_other = 1; wd.set[0].n = (int32_t)ra.n;
    wd.set[1].rs_self = 1; wd.set[1].rs_other = 0; wd.set[1].n = (int32_t)rb.n;
    WaveIO io;
    io.shard = true;
    int tile_rows[2] = {g.tile_row_begin, g.tile_row_end};
    io.out_tile_rows = tile_rows;
    if (balance_world > 0) { io.balance_rank = balance_rank; io.balance_world = balance_world; }
    Slot &sl = c->S();
    OKR(run_wave(*c, sl, st, p, g, wd, io));
    g.tile_row_begin = tile_rows[0];
    g.tile_row_end = tile_rows[1];
    row_begin = std::min(p.R, g.tile_row_begin * kTile);
    row_end = std::min(p.R, g.tile_row_end * kTile);
    if (out_rows) { out_rows[0] = row_begin; out_rows[1] = row_end; }
    OKR(c->S().misc.ensure(64));
    const int nparts = g.NT * g.NT;
    const int lo = g.tile_row_begin * g.NT, cnt = (g.tile_row_end - g.tile_row_begin) * g.NT;
    {   // local normaliser partials: tiles [lo, lo+cnt) of each set (set 1 sits nparts after set 0)
        KTimer t(*c, 5, st);
        k_norm_totals<<<1, kJsThreads, 0, st>>>(cnt, nparts, sl.tilesum.as<double>() + lo, c->S().misc.as<double>());
    }
    OKR(ensure_host_results(*c, 4));
    CU(cudaMemcpyAsync(c->h_js, c->S().misc.p, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_status, &sl.rstat.as<RawStat>()[0].flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_status + 1, &sl.rstat.as<RawStat>()[1].flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    release_scratch(*c, st, true);
    if ((c->h_status[0] | c->h_status[1]) & 1)
        return fail(KDEJS_ERR_INF, "Input X contains infinity or a value too large for dtype('float64').");
    out_norm[0] = c->h_js[0];
    out_norm[1] = c->h_js[1];
    c->shard_wd = wd; c->shard_geom = g;
    c->shard_kb = row_begin * p.R;
    c->shard_ke = row_end * p.R;
    c->shard_valid = true;
    c->shard_slot = c->cur;
    return KDEJS_OK;
}

static int shard_check(const Params &p, const double *a, int64_t na, const double *b, int64_t nb, int row_begin,
                       int row_end, const double *out_norm) {
    OKR(check_params(p));
    OKR(check_set(a, na, "df1"));
    OKR(check_set(b, nb, "df2"));
    if (p.algo != KDEJS_ALGO_BINNED && p.algo != KDEJS_ALGO_BINNED32)
        return fail(KDEJS_ERR_ARG, "row sharding is implemented for the binned algorithms");
    if (row_begin < 0 || row_end > p.R || row_begin >= row_end) return fail(KDEJS_ERR_ARG, "bad row range");
    if (row_begin % kTile != 0 || (row_end % kTile != 0 && row_end != p.R))
        return fail(KDEJS_ERR_ARG, "row range must be aligned to 4 rows (or end at the last row)");
    if (!out_norm) return fail(KDEJS_ERR_ARG, "out_norm is null");
    return KDEJS_OK;
}

int kdejs_shard_phase1(int device, const double *a, int64_t na, const double *b, int64_t nb, int resolution,
                       double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                       int row_begin, int row_end, double out_norm[2]) {
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(shard_check(p, a, na, b, nb, row_begin, row_end, out_norm));
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t offb = ((size_t)na * 16 + 255) / 256 * 256;
    OKR(c->pool_raw.ensure(offb + (size_t)nb * 16));
    OKR(h2d_ranges(device, {Stager::Range{c->pool_raw.p, a, (size_t)na * 16},
                            Stager::Range{(char *)c->pool_raw.p + offb, b, (size_t)nb * 16}}, c->s_compute));
    return shard_phase1_core(c, p, RawSet{(const double *)c->pool_raw.p, na},
                             RawSet{(const double *)((char *)c->pool_raw.p + offb), nb}, row_begin, row_end, out_norm);
}

int kdejs_shard_phase1_dev(int device, const double *a_dev, int64_t na, const double *b_dev, int64_t nb,
                           int resolution, double bandwidth, double gamma, double eps, int log_flag, int kernel,
                           int algo, int row_begin, int row_end, double out_norm[2]) {
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(shard_check(p, a_dev, na, b_dev, nb, row_begin, row_end, out_norm));
    if ((((uintptr_t)a_dev) | ((uintptr_t)b_dev)) & 15) return fail(KDEJS_ERR_ARG, "device point sets must be 16-byte aligned");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    return shard_phase1_core(c, p, RawSet{a_dev, na}, RawSet{b_dev, nb}, row_begin, row_end, out_norm);
}

int kdejs_shard_phase1_balanced(int device, const double *a, int64_t na, const double *b, int64_t nb,
                                int on_device, int resolution, double bandwidth, double gamma, double eps,
                                int log_flag, int kernel, int algo, int rank, int world, double out_norm[2],
                                int out_rows[2]) {
    Params p{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(shard_check(p, a, na, b, nb, 0, resolution, out_norm));
    if (world < 1 || rank < 0 || rank >= world || !out_rows) return fail(KDEJS_ERR_ARG, "bad rank/world");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    RawSet ra{a, na}, rb{b, nb};
    if (!on_device) {
        const size_t offb = ((size_t)na * 16 + 255) / 256 * 256;
        OKR(c->pool_raw.ensure(offb + (size_t)nb * 16));
        OKR(h2d_ranges(device, {Stager::Range{c->pool_raw.p, a, (size_t)na * 16},
                                Stager::Range{(char *)c->pool_raw.p + offb, b, (size_t)nb * 16}}, c->s_compute));
        ra.p = (const double *)c->pool_raw.p;
        rb.p = (const double *)((char *)c->pool_raw.p + offb);
    } else if ((((uintptr_t)a) | ((uintptr_t)b)) & 15) {
        return fail(KDEJS_ERR_ARG, "device point sets must be 16-byte aligned");
    }
    return shard_phase1_core(c, p, ra, rb, 0, resolution, out_norm, rank, world, out_rows);
}

int kdejs_shard_phase2(int device, const double total_norm[2], double out_lr[2]) {
    if (!total_norm || !out_lr) return fail(KDEJS_ERR_ARG, "null argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->shard_valid) return fail(KDEJS_ERR_ARG, "kdejs_shard_phase2 called without a successful phase 1");
    cudaStream_t st = c->s_compute;
    OKR(acquire_scratch(*c, st, c->shard_slot));
    Slot &sl = c->S();
    const GridGeom &g = c->shard_geom;
    const int G = g.R * g.R;
    double *d = c->S().misc.as<double>();      // [0,1] totals in, [2,3] (left,right) out
    CU(cudaMemcpyAsync(d, total_norm, 16, cudaMemcpyHostToDevice, st));
    const int n = c->shard_ke - c->shard_kb;
    if (n <= 0) { out_lr[0] = 0.0; out_lr[1] = 0.0; return KDEJS_OK; }     // this rank owns no rows
    const int js_blocks = (n + kJsNodesPerBlock - 1) / kJsNodesPerBlock;
    OKR(sl.jspart.ensure(sizeof(double) * 2 * (size_t)js_blocks));
    {
        KTimer t(*c, 5, st);
        k_js<<<dim3(js_blocks, 1), kJsThreads, 0, st>>>(
            c->shard_wd, G, g.NT * g.NT, c->shard_kb, c->shard_ke, sl.pz.as<double>(), d,
            sl.rstat.as<RawStat>(), nullptr, sl.jspart.as<double>(),
            sl.tickets.as<unsigned>() + kMaxWaveSets, nullptr, nullptr, d + 2, nullptr, 0);
    }
    CU(cudaGetLastError());
    OKR(ensure_host_results(*c, 4));
    CU(cudaMemcpyAsync(c->h_js, d + 2, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    release_scratch(*c, st, true);
    out_lr[0] = c->h_js[0];
    out_lr[1] = c->h_js[1];
    return KDEJS_OK;
}


// ============================================================================================
// Band mode: one huge evaluation over several GPUs with the POINTS sharded too.
// Layout of a rank's count block (u32): cellstart[2][C*C + 1] of its locally sorted slices, then
// rowstart[2][C + 1] (= cellstart at the first cell of every cell row), then kBandHandleWords words the caller may
// fill with the CUDA IPC handle of the buffer its sorted points live in (all zero: none) -- the handle travels with
// the counts, so a reallocated buffer needs no collective of its own.
namespace {
constexpr int kBandHandleWords = 16;
static_assert(sizeof(cudaIpcMemHandle_t) == 4 * kBandHandleWords, "a CUDA IPC handle is 64 bytes");
struct BandDims { int C, NT, ncells; size_t block_u32; };
BandDims band_dims(const GridGeom &g) {
    BandDims d;
    d.C = g.C; d.NT = g.NT; d.ncells = g.C * g.C;
    d.block_u32 = 2 * (size_t)(d.ncells + 1) + 2 * (size_t)(g.C + 1) + kBandHandleWords;
    return d;
}
int band_params(int resolution, double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo, Params &p) {
    p = Params{resolution, bandwidth, gamma, eps, log_flag ? 1 : 0, kernel, algo};
    OKR(check_params(p));
    if (p.algo != KDEJS_ALGO_BINNED && p.algo != KDEJS_ALGO_BINNED32)
        return fail(KDEJS_ERR_ARG, "band mode is implemented for the binned algorithms");
    return KDEJS_OK;
}
}  // namespace

int kdejs_band_geometry(int resolution, double bandwidth, int kernel, int out[4]) {
    if (!out) return fail(KDEJS_ERR_ARG, "out is null");
    Params p{resolution, bandwidth, 1.0, 0.0, 0, kernel, KDEJS_ALGO_BINNED};
    OKR(check_params(p));
    const GridGeom g = make_geom(p);
    const BandDims d = band_dims(g);
    out[0] = d.C; out[1] = d.NT; out[2] = (int)d.block_u32; out[3] = g.use_moments;
    return KDEJS_OK;
}

int kdejs_band_minmax_dev(int device, const double *a_dev, int64_t na, const double *b_dev, int64_t nb,
                          int log_flag, double eps, double *out_stats_dev, void *stream) {
    OKR(check_set(a_dev, na, "df1 slice"));
    OKR(check_set(b_dev, nb, "df2 slice"));
    if (!out_stats_dev) return fail(KDEJS_ERR_ARG, "out_stats_dev is null");
    if ((((uintptr_t)a_dev) | ((uintptr_t)b_dev)) & 15) return fail(KDEJS_ERR_ARG, "device point sets must be 16-byte aligned");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    OKR(acquire_scratch(*c, st));
    Slot &sl = c->S();
    WaveDesc wd;
    memset(&wd, 0, sizeof wd);
    wd.raw[0] = RawSet{a_dev, na};
    wd.raw[1] = RawSet{b_dev, nb};
    wd.n_raw = 2;
    const int mm_blocks = std::max(8, std::min(64, 4 * c->sm_count / 2));
    OKR(sl.rpart.ensure(sizeof(RawStat) * kMaxWaveSets * 64));
    static_assert(sizeof(RawStat) == 40, "RawStat travels as 5 doubles");
    {
        KTimer t(*c, 0, st);
        k_minmax<<<dim3(mm_blocks, 2), kMinmaxThreads, 0, st>>>(wd, log_flag ? 1 : 0, eps, sl.rpart.as<RawStat>(),
                                                               sl.tickets.as<unsigned>(),
                                                               reinterpret_cast<RawStat *>(out_stats_dev));
    }
    CU(cudaGetLastError());
    return release_scratch(*c, st, false);
}

int kdejs_band_sort_dev(int device, const double *a_dev, int64_t na, const double *b_dev, int64_t nb, int resolution,
                        double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                        const double *stats_all_dev, int world, double *out_sorted_dev, uint32_t *out_block_dev,
                        void *stream) {
    Params p;
    OKR(band_params(resolution, bandwidth, gamma, eps, log_flag, kernel, algo, p));
    OKR(check_set(a_dev, na, "df1 slice"));
    OKR(check_set(b_dev, nb, "df2 slice"));
    if (!stats_all_dev || !out_sorted_dev || !out_block_dev || world < 1) return fail(KDEJS_ERR_ARG, "null argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    OKR(acquire_scratch(*c, st));
    Slot &sl = c->S();
    GridGeom g = make_geom(p);
    const BandDims d = band_dims(g);
    OKR(sl.rstat.ensure(sizeof(RawStat) * kMaxWaveSets));
    k_band_fold_stats<<<1, 32, 0, st>>>(reinterpret_cast<const RawStat *>(stats_all_dev), world, sl.rstat.as<RawStat>());
    c->launches++;
    WaveDesc wd;
    memset(&wd, 0, sizeof wd);
    wd.raw[0] = RawSet{a_dev, na};
    wd.raw[1] = RawSet{b_dev, nb};
    wd.n_raw = 2; wd.n_sets = 2;
    wd.set[0].rs_self = 0; wd.set[0].rs_other = 1; wd.set[0].n = (int32_t)na;
    wd.set[1].rs_self = 1; wd.set[1].rs_other = 0; wd.set[1].n = (int32_t)nb;
    WaveIO io;
    io.shard = true; io.rstat_ready = true; io.sort_only = true;
    Params ps = p;
    ps.algo = KDEJS_ALGO_BINNED;                 // the sender needs no fixed-point copy
    OKR(run_wave(*c, sl, st, ps, g, wd, io));
    CU(cudaMemcpyAsync(out_sorted_dev, sl.sorted.p, (size_t)(na + nb) * 16, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(out_block_dev, sl.cellstart.p, 2 * (size_t)(d.ncells + 1) * 4, cudaMemcpyDeviceToDevice, st));
    k_band_rowstarts<<<2, 128, 0, st>>>(sl.cellstart.as<uint32_t>(), d.C, out_block_dev + 2 * (size_t)(d.ncells + 1));
    c->launches++;
    CU(cudaGetLastError());
    return release_scratch(*c, st, false);
}

int kdejs_band_cuts(const double *colwork, int nt, int world, int *cuts) {
    if (!colwork || !cuts || nt < 1 || world < 1) return fail(KDEJS_ERR_ARG, "bad argument");
    double total = 0.0;
    for (int t = 0; t < nt; ++t) total += colwork[t];
    for (int r = 0; r <= world; ++r) cuts[r] = nt;
    cuts[0] = 0;
    double run = 0.0;
    int r = 1;
    for (int t = 0; t < nt && r < world; ++t) {
        run += colwork[t];
        while (r < world && run >= total * r / world) cuts[r++] = t + 1;
    }
    return KDEJS_OK;
}

int kdejs_band_cell_rows(int resolution, double bandwidth, int ty_begin, int ty_end, int out_cy[2]) {
    if (!out_cy) return fail(KDEJS_ERR_ARG, "out_cy is null");
    Params p{resolution, bandwidth, 1.0, 0.0, 0, KDEJS_KERNEL_EPANECHNIKOV, KDEJS_ALGO_BINNED};
    OKR(check_params(p));
    const GridGeom g = make_geom(p);
    if (ty_begin < 0 || ty_end > g.NT || ty_begin > ty_end) return fail(KDEJS_ERR_ARG, "bad tile-column range");
    if (ty_begin == ty_end) { out_cy[0] = 0; out_cy[1] = -1; return KDEJS_OK; }      // empty band: no rows
    // the arithmetic of tile_cells() for the band's first and last tile column (cell_of is monotone)
    auto cell_of_h = [&](double x) {
        const double v = std::floor((x - g.cell_x0) * g.cell_invw);
        return (int)std::min<double>(std::max<double>(v, 0.0), (double)(g.C - 1));
    };
    const double y_lo = (double)(ty_begin * kTile) * g.step + g.gmin;
    const double y_hi = (double)std::min(ty_end * kTile - 1, g.R - 1) * g.step + g.gmin;
    out_cy[0] = cell_of_h(y_lo - g.h - g.margin);
    out_cy[1] = cell_of_h(y_hi + g.h + g.margin);
    return KDEJS_OK;
}

int kdejs_band_layout(const double *colwork, int nt, int world, int resolution, double bandwidth, int *out_cuts,
                      int *out_rows) {
    if (!out_rows) return fail(KDEJS_ERR_ARG, "out_rows is null");
    OKR(kdejs_band_cuts(colwork, nt, world, out_cuts));
    for (int r = 0; r < world; ++r) OKR(kdejs_band_cell_rows(resolution, bandwidth, out_cuts[r], out_cuts[r + 1], out_rows + 2 * r));
    return KDEJS_OK;
}

int kdejs_band_plan(int device, const uint32_t *blocks_all_dev, int world, int resolution, double bandwidth,
                    double gamma, double eps, int log_flag, int kernel, int algo, double *out_colwork,
                    uint32_t *out_rowstart, int *out_use_moments, void *stream) {
    Params p;
    OKR(band_params(resolution, bandwidth, gamma, eps, log_flag, kernel, algo, p));
    if (!blocks_all_dev || !out_colwork || !out_rowstart || !out_use_moments || world < 1)
        return fail(KDEJS_ERR_ARG, "null argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    OKR(acquire_scratch(*c, st));
    Slot &sl = c->S();
    GridGeom g = make_geom(p);
    const BandDims d = band_dims(g);
    // the global cell table (the numbers a single-GPU sort produces), the work per tile column from it, and every
    // rank's row starts (world x 2 x (C + 1) u32, a strided copy): one synchronisation for all of it
    OKR(sl.cellstart.ensure(2 * (size_t)(d.ncells + 1) * sizeof(uint32_t)));
    OKR(sl.misc.ensure(64 + sizeof(double) * (size_t)(g.NT + 1)));
    OKR(sl.bandtot.ensure(2 * (size_t)d.ncells * sizeof(uint32_t)));
    k_band_cell_totals<<<dim3((d.ncells + 255) / 256, 2), 256, 0, st>>>(blocks_all_dev, world, d.block_u32, d.ncells, 0,
                                                                        d.ncells, sl.bandtot.as<uint32_t>());
    k_band_cells<<<2, 1024, 0, st>>>(sl.bandtot.as<uint32_t>(), d.ncells, sl.cellstart.as<uint32_t>());
    c->launches++;
    WaveDesc wd;
    memset(&wd, 0, sizeof wd);
    wd.n_sets = 2;
    double *d_cw = sl.misc.as<double>() + 8;
    // moment cells: the single-GPU rule applied to the WHOLE evaluation, so every rank (and the one-GPU run) agrees
    // one (set, tile) per thread: the work numbers are integers, so the block's sum does not depend on its shape
    k_tile_colwork<<<g.NT, 512, 0, st>>>(wd, g, sl.cellstart.as<uint32_t>(), d_cw, (long long)128 * 2 * d.ncells);
    c->launches += 2;
    CU(cudaGetLastError());
    const size_t rs_elems = 2 * (size_t)(d.C + 1) + kBandHandleWords;     // row starts and the handle words behind them
    std::vector<double> cw((size_t)g.NT + 1);
    CU(cudaMemcpy2DAsync(out_rowstart, rs_elems * 4, blocks_all_dev + 2 * (size_t)(d.ncells + 1), d.block_u32 * 4,
                         rs_elems * 4, (size_t)world, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cw.data(), d_cw, sizeof(double) * cw.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(out_colwork, cw.data(), sizeof(double) * (size_t)g.NT);
    *out_use_moments = cw[(size_t)g.NT] != 0.0 ? 1 : 0;
    return release_scratch(*c, st, true);
}

}  // extern "C"
namespace {
// peer = true: the sources are the ranks' own sorted arrays (Ctx::peer_map), stage_* unused
int band_phase1_core(int device, bool peer, const double *stage_a_dev, int64_t na, const double *stage_b_dev, int64_t nb,
                     int64_t na_total, int64_t nb_total, const uint32_t *blocks_all_dev, int world,
                     const double *stats_all_dev, int resolution, double bandwidth, double gamma, double eps,
                     int log_flag, int kernel, int algo, int ty_begin, int ty_end, int cy_lo, int cy_hi,
                     int use_moments, double *out_norm_dev, void *stream) {
    Params p;
    OKR(band_params(resolution, bandwidth, gamma, eps, log_flag, kernel, algo, p));
    if (!blocks_all_dev || !stats_all_dev || !out_norm_dev || world < 1) return fail(KDEJS_ERR_ARG, "null argument");
    if (na < 0 || nb < 0 || na_total < 1 || nb_total < 1 || (!peer && ((na > 0 && !stage_a_dev) || (nb > 0 && !stage_b_dev))))
        return fail(KDEJS_ERR_ARG, "bad staging buffers");
    if (na + nb > 0x7fffffffLL / 4) return fail(KDEJS_ERR_ARG, "too many points for one band");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    PeerPtrs peers;
    memset(&peers, 0, sizeof peers);
    if (peer) {
        if (c->peer_world != world) return fail(KDEJS_ERR_ARG, "kdejs_band_peer_open was not called for this world size");
        for (int r = 0; r < world; ++r) {
            peers.p[r] = reinterpret_cast<const double2 *>(r == c->peer_rank ? c->peer_buf.p : c->peer_map[r].ptr);
            if (!peers.p[r]) return fail(KDEJS_ERR_ARG, "a peer's sorted points are not mapped");
        }
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    OKR(acquire_scratch(*c, st));
    c->band_valid = false;
    Slot &sl = c->S();
    GridGeom g = make_geom(p);
    const BandDims d = band_dims(g);
    if (ty_begin < 0 || ty_end > g.NT || ty_begin > ty_end) return fail(KDEJS_ERR_ARG, "bad tile-column range");
    if (ty_begin < ty_end && (cy_lo < 0 || cy_hi >= d.C || cy_lo > cy_hi)) return fail(KDEJS_ERR_ARG, "bad cell-row range");
    g.tile_col_begin = ty_begin;
    g.tile_col_end = ty_end;
    OKR(sl.rstat.ensure(sizeof(RawStat) * kMaxWaveSets));
    OKR(sl.tilesum.ensure(2 * (size_t)g.NT * g.NT * sizeof(double)));
    OKR(c->S().misc.ensure(64));
    k_band_fold_stats<<<1, 32, 0, st>>>(reinterpret_cast<const RawStat *>(stats_all_dev), world, sl.rstat.as<RawStat>());
    c->launches++;
    if (ty_begin == ty_end) {                       // this rank owns no tile column: zero partial normalisers
        CU(cudaMemsetAsync(out_norm_dev, 0, 16, st));
        c->band_geom = g; c->band_ty0 = ty_begin; c->band_ty1 = ty_end; c->band_slot = c->cur; c->band_valid = true;
        return release_scratch(*c, st, false);
    }
    const bool fast32 = (p.algo == KDEJS_ALGO_BINNED32) && g.q_shift > 0;
    const int64_t P = na + nb;
    OKR(sl.sorted.ensure((size_t)std::max<int64_t>(P, 1) * sizeof(double2)));
    if (fast32 && !kFp32FromDoubles) OKR(sl.sorted_q.ensure((size_t)std::max<int64_t>(P, 1) * sizeof(uint2)));
    OKR(sl.cellstart.ensure(2 * (size_t)(d.ncells + 1) * sizeof(uint32_t)));
    const int c_lo = cy_lo * d.C, c_hi = (cy_hi + 1) * d.C;
    OKR(sl.bandtot.ensure(2 * (size_t)d.ncells * sizeof(uint32_t)));
    k_band_cell_totals<<<dim3((d.ncells + 255) / 256, 2), 256, 0, st>>>(blocks_all_dev, world, d.block_u32, d.ncells, c_lo,
                                                                        c_hi, sl.bandtot.as<uint32_t>());
    k_band_cells<<<2, 1024, 0, st>>>(sl.bandtot.as<uint32_t>(), d.ncells, sl.cellstart.as<uint32_t>());
    c->launches++;
    const dim3 merge_grid((unsigned)((std::max<int64_t>(std::max(na, nb), 1) + 256 * kMergePer - 1) / (256 * kMergePer)), 2);
    if (peer)
        k_band_merge<true><<<merge_grid, 256, 0, st>>>(
            nullptr, nullptr, peers, blocks_all_dev,
            world, d.block_u32, g, cy_lo, cy_hi, sl.cellstart.as<uint32_t>(), na, sl.sorted.as<double2>(),
            fast32 && !kFp32FromDoubles ? sl.sorted_q.as<uint2>() : nullptr);
    else
        k_band_merge<false><<<merge_grid, 256, 0, st>>>(
            reinterpret_cast<const double2 *>(stage_a_dev), reinterpret_cast<const double2 *>(stage_b_dev), peers, blocks_all_dev,
            world, d.block_u32, g, cy_lo, cy_hi, sl.cellstart.as<uint32_t>(), na, sl.sorted.as<double2>(),
            fast32 && !kFp32FromDoubles ? sl.sorted_q.as<uint2>() : nullptr);
    c->launches += 2;
    WaveDesc wd;
    memset(&wd, 0, sizeof wd);
    wd.n_raw = 2; wd.n_sets = 2;
    wd.set[0].rs_self = 0; wd.set[0].rs_other = 1; wd.set[0].n = (int32_t)na;
    wd.set[1].rs_self = 1; wd.set[1].rs_other = 0; wd.set[1].n = (int32_t)nb;
    WaveIO io;
    io.shard = true; io.rstat_ready = true; io.presorted = true; io.identity_scaler = true;
    io.force_moments = use_moments ? 1 : 0;
    const int64_t ng[2] = {na_total, nb_total};
    io.n_global = ng;
    OKR(run_wave(*c, sl, st, p, g, wd, io));
    {
        KTimer t(*c, 5, st);
        k_band_norm_totals<<<1, kJsThreads, 0, st>>>(sl.tilesum.as<double>(), g.NT, ty_begin, ty_end, out_norm_dev);
    }
    CU(cudaGetLastError());
    c->band_geom = g; c->band_ty0 = ty_begin; c->band_ty1 = ty_end; c->band_slot = c->cur; c->band_valid = true;
    return release_scratch(*c, st, false);
}
}  // namespace
extern "C" {

int kdejs_band_phase1_dev(int device, const double *stage_a_dev, int64_t na, const double *stage_b_dev, int64_t nb,
                          int64_t na_total, int64_t nb_total, const uint32_t *blocks_all_dev, int world,
                          const double *stats_all_dev, int resolution, double bandwidth, double gamma, double eps,
                          int log_flag, int kernel, int algo, int ty_begin, int ty_end, int cy_lo, int cy_hi,
                          int use_moments, double *out_norm_dev, void *stream) {
    return band_phase1_core(device, false, stage_a_dev, na, stage_b_dev, nb, na_total, nb_total, blocks_all_dev, world,
                            stats_all_dev, resolution, bandwidth, gamma, eps, log_flag, kernel, algo, ty_begin, ty_end,
                            cy_lo, cy_hi, use_moments, out_norm_dev, stream);
}

int kdejs_band_phase1_peer_dev(int device, int64_t na, int64_t nb, int64_t na_total, int64_t nb_total,
                               const uint32_t *blocks_all_dev, int world, const double *stats_all_dev, int resolution,
                               double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                               int ty_begin, int ty_end, int cy_lo, int cy_hi, int use_moments, double *out_norm_dev,
                               void *stream) {
    return band_phase1_core(device, true, nullptr, na, nullptr, nb, na_total, nb_total, blocks_all_dev, world,
                            stats_all_dev, resolution, bandwidth, gamma, eps, log_flag, kernel, algo, ty_begin, ty_end,
                            cy_lo, cy_hi, use_moments, out_norm_dev, stream);
}

int kdejs_band_peer_buffer(int device, int64_t points, void **out_ptr, uint32_t *out_handle_words) {
    if (points < 0 || !out_ptr || !out_handle_words) return fail(KDEJS_ERR_ARG, "bad argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t need = (size_t)std::max<int64_t>(points, 1) * sizeof(double2);
    if (need > c->peer_buf.cap) {
        // Grow with headroom: every reallocation makes all peers map the buffer again.  An exported allocation
        // must not be freed while another process has it mapped, and peers drop their mapping only when they see
        // the new handle: the old buffer is parked until the context goes away.
        if (c->peer_buf.p) c->peer_retired.push_back(c->peer_buf.p);
        c->peer_buf.p = nullptr; c->peer_buf.cap = 0;
        c->peer_handle_valid = false;
        OKR(c->peer_buf.ensure(need + need / 4));
    }
    if (!c->peer_handle_valid) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->peer_buf.p));
        memcpy(c->peer_handle, &h, 64);
        c->peer_handle_valid = true;
    }
    *out_ptr = c->peer_buf.p;
    memcpy(out_handle_words, c->peer_handle, 64);
    return KDEJS_OK;
}

int kdejs_band_peer_open(int device, const uint32_t *handle_words, int world, int rank) {
    if (!handle_words || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(KDEJS_ERR_ARG, "bad argument (at most 16 ranks share points over peer memory)");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    static const unsigned char zero[64] = {0};
    for (int r = 0; r < world; ++r) {
        const unsigned char *h = reinterpret_cast<const unsigned char *>(handle_words) + (size_t)r * 64;
        if (memcmp(h, zero, 64) == 0) return fail(KDEJS_ERR_ARG, "a rank exported no peer buffer");
        Ctx::PeerMap &m = c->peer_map[r];
        if (r == rank) {
            if (!c->peer_handle_valid || memcmp(h, c->peer_handle, 64) != 0)
                return fail(KDEJS_ERR_ARG, "this rank's handle is not the one kdejs_band_peer_buffer returned");
            continue;
        }
        if (m.open && memcmp(m.handle, h, 64) == 0) continue;           // still the mapping of the last evaluation
        if (m.open) {
            CU(cudaDeviceSynchronize());                                 // nothing of ours may still be reading it
            CU(cudaIpcCloseMemHandle(m.ptr));
            m.open = false; m.ptr = nullptr;
        }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, h, 64);
        void *ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
        memcpy(m.handle, h, 64);
        m.ptr = ptr;
        m.open = true;
    }
    for (int r = world; r < kMaxPeers; ++r) {
        Ctx::PeerMap &m = c->peer_map[r];
        if (m.open) { CU(cudaDeviceSynchronize()); cudaIpcCloseMemHandle(m.ptr); m.open = false; m.ptr = nullptr; }
    }
    c->peer_world = world;
    c->peer_rank = rank;
    return KDEJS_OK;
}

int kdejs_band_phase2_dev(int device, const double *norms_all_dev, int world, double *out_lr_dev, void *stream) {
    if (!norms_all_dev || !out_lr_dev || world < 1) return fail(KDEJS_ERR_ARG, "null argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->band_valid) return fail(KDEJS_ERR_ARG, "kdejs_band_phase2_dev called without a successful phase 1");
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    OKR(acquire_scratch(*c, st, c->band_slot));
    Slot &sl = c->S();
    const GridGeom &g = c->band_geom;
    double *d_tot = c->S().misc.as<double>();
    k_band_fold_norms<<<1, 32, 0, st>>>(norms_all_dev, world, d_tot);
    c->launches++;
    if (c->band_ty0 == c->band_ty1) {
        // no tiles: zero partial sums, but the infinity flag of the inputs still has to travel
        CU(cudaMemsetAsync(out_lr_dev, 0, 24, st));
        return release_scratch(*c, st, false);
    }
    OKR(sl.jspart.ensure(sizeof(double) * 2 * (size_t)g.NT));
    {
        KTimer t(*c, 5, st);
        k_band_js<<<g.NT, kJsThreads, 0, st>>>(g, c->band_ty0, c->band_ty1, sl.pz.as<double>(), d_tot,
                                              sl.rstat.as<RawStat>(), sl.jspart.as<double>(),
                                              sl.tickets.as<unsigned>() + kMaxWaveSets, out_lr_dev);
    }
    CU(cudaGetLastError());
    return release_scratch(*c, st, false);
}

int kdejs_profile_enable(int device, int on) {
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(resolve_pending(*c));
    c->prof_on = on != 0;
    c->prof_mode = on;
    return KDEJS_OK;
}

int kdejs_profile_read(int device, int which, double *ms_total, int64_t *launches) {
    if (which < 0 || which >= KDEJS_NUM_KERNELS) return fail(KDEJS_ERR_ARG, "bad kernel index");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(resolve_pending(*c));
    if (ms_total) *ms_total = c->prof_ms[which];
    if (launches) *launches = c->prof_n[which];
    return KDEJS_OK;
}

int kdejs_profile_reset(int device) {
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(resolve_pending(*c));
    for (int i = 0; i < KDEJS_NUM_KERNELS; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    return KDEJS_OK;
}

int kdejs_launch_count(int device, int64_t *count, int reset) {
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) *count = c->launches;
    if (reset) c->launches = 0;
    return KDEJS_OK;
}

int kdejs_work_counters(int device, double *pair_tests, int reset) {
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaDeviceSynchronize());
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, c->d_work, sizeof v, cudaMemcpyDeviceToHost));
    if (pair_tests) *pair_tests = (double)v;
    if (reset) CU(cudaMemset(c->d_work, 0, sizeof v));
    return KDEJS_OK;
}

int kdejs_fp32_stats(int device, double out[5], int reset) {
    if (!out) return fail(KDEJS_ERR_ARG, "out is null");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaDeviceSynchronize());
    unsigned long long v[8];
    CU(cudaMemcpy(v, c->d_stats32, sizeof v, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; ++i) out[i] = (double)v[i];
    if (reset) CU(cudaMemset(c->d_stats32, 0, sizeof v));
    return KDEJS_OK;
}

int kdejs_loop_peak(int device, int which, double *pair_tests_per_s) {
    if (which < 0 || which > 2 || !pair_tests_per_s) return fail(KDEJS_ERR_ARG, "bad argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(acquire_scratch(*c, c->s_compute));
    OKR(c->S().misc.ensure(64 + 1024 * 16 + 1024 * 8));
    double *sink = c->S().misc.as<double>();
    double2 *pts = reinterpret_cast<double2 *>(c->S().misc.as<char>() + 64);
    uint2 *ptsq = reinterpret_cast<uint2 *>(c->S().misc.as<char>() + 64 + 1024 * 16);
    std::vector<double2> hp(1024);
    std::vector<uint2> hq(1024);
    for (int i = 0; i < 1024; ++i) {
        const double x = 0.2 + 0.1 * ((i * 37) % 1024) / 1024.0, y = 0.2 + 0.1 * ((i * 101) % 1024) / 1024.0;
        hp[i] = make_double2(x, y);
        hq[i] = make_uint2(0x10000000u + (uint32_t)((x - 0.25) * 33.0 * 536870912.0), 0x10000000u + (uint32_t)((y - 0.25) * 33.0 * 536870912.0));
    }
    CU(cudaMemcpyAsync(pts, hp.data(), 1024 * 16, cudaMemcpyHostToDevice, c->s_compute));
    CU(cudaMemcpyAsync(ptsq, hq.data(), 1024 * 8, cudaMemcpyHostToDevice, c->s_compute));
    const int blocks = c->sm_count * 5, trips = 1 << 13;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, c->s_compute));
        if (which == 0) k_loop_peak<0><<<blocks, kDensThreads, 0, c->s_compute>>>(trips, pts, ptsq, sink);
        else if (which == 1) k_loop_peak<1><<<blocks, kDensThreads, 0, c->s_compute>>>(trips, pts, ptsq, sink);
        else k_loop_peak<2><<<blocks, kDensThreads, 0, c->s_compute>>>(trips, pts, ptsq, sink);
        CU(cudaEventRecord(e1, c->s_compute));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    release_scratch(*c, c->s_compute, true);
    *pair_tests_per_s = (double)blocks * kDensThreads * (double)trips * 16.0 / (best * 1e-3);
    return KDEJS_OK;
}

int kdejs_peak_flops(int device, int which, double *tflops, double *sm_clock_mhz) {
    if (which < 0 || which > 4 || !tflops) return fail(KDEJS_ERR_ARG, "bad argument");
    Ctx *c;
    OKR(get_ctx(device, &c));
    std::lock_guard<std::mutex> lk(c->mu);
    OKR(c->S().misc.ensure(64));
    const int blocks = c->sm_count * 8, iters = 1 << 15;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, c->s_compute));
        if (which == 0) k_peak<0><<<blocks, 256, 0, c->s_compute>>>(iters, c->S().misc.as<float>(), c->S().misc.as<double>());
        else if (which == 1) k_peak<1><<<blocks, 256, 0, c->s_compute>>>(iters, c->S().misc.as<float>(), c->S().misc.as<double>());
        else if (which == 2) k_peak<2><<<blocks, 256, 0, c->s_compute>>>(iters, c->S().misc.as<float>(), c->S().misc.as<double>());
        else if (which == 3) k_peak<3><<<blocks, 256, 0, c->s_compute>>>(iters, c->S().misc.as<float>(), c->S().misc.as<double>());
        else k_peak<4><<<blocks, 256, 0, c->s_compute>>>(iters, c->S().misc.as<float>(), c->S().misc.as<double>());
        CU(cudaEventRecord(e1, c->s_compute));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double fmas = (double)blocks * 256.0 * iters * 8.0 * (which == 2 ? 2.0 : 1.0);
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *sm_clock_mhz = khz / 1000.0;
    }
    return KDEJS_OK;
}

}  // extern "C"
