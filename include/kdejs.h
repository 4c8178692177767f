/*
 * kdejs.h -- C ABI of libkdejs.so, the B200 (sm_100a) implementation of PopPUNK-mod's
 * KDE discrepancy.  Plain pointers and sizes only; no torch / CUDA types cross this line
 * (a CUDA stream is passed as an opaque void*).
 *
 * The reference has no FFI of its own: its hot path is Python calling compiled wheels
 * (sklearn Cython / scipy C).  Each entry point below replaces the Python-level interface
 * named beside it (paths relative to the reference repository); INTEGRATION.md shows the
 * ctypes binding a maintainer would add to KDE_distance.py.
 *
 * Conventions
 *   - point sets are row-major (N,2) float64, column 0 = core distance, 1 = accessory
 *   - grids are flat (R*R) float64, node k = ix*R + iy at (lin[ix], lin[iy]),
 *     lin = linspace(gmin, gmax, R): the order of get_grid()'s `xy` (KDE_distance.py:17-23)
 *   - every function returns KDEJS_OK or an error code; kdejs_last_error() gives the
 *     message for the calling thread.  There is no CPU fallback: without a usable
 *     sm_100 device every compute entry point fails with KDEJS_ERR_CUDA.
 *   - the caller owns all buffers it passes; the library owns its device scratch (grown
 *     on demand, cached per device, released by kdejs_shutdown).  No pointer is retained
 *     after a call returns, except by the *_dev entry points until their stream work ends.
 *   - nothing touches CUDA at load time (fork-safe: ELFI's multiprocessing client and the
 *     scripts' Pools fork after import); a context is created on first use per process.
 */
#ifndef KDEJS_H
#define KDEJS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDEJS_VERSION 100

/* status codes */
#define KDEJS_OK 0
#define KDEJS_ERR_CUDA 1     /* CUDA runtime failure or no device            -> RuntimeError */
#define KDEJS_ERR_ARG 2      /* bad shape / size / parameter                 -> ValueError   */
#define KDEJS_ERR_INF 3      /* input contains infinity (sklearn ValueError) -> ValueError   */
#define KDEJS_ERR_CAPACITY 4 /* kdejs_tsv_load_into: the caller's buffer is too small (rows needed are reported) */

/* kernel (sklearn KernelDensity(kernel=...), KDE_distance.py:50-51) */
#define KDEJS_KERNEL_EPANECHNIKOV 0   /* the reference's kernel                            */
#define KDEJS_KERNEL_GAUSSIAN 1       /* extra mode; dense algorithm only */

/* algorithm for the density sum (results agree within the documented tolerances) */
#define KDEJS_ALGO_BINNED 0   /* fp64, points counting-sorted into cells, only in-reach   */
                              /* (tile, cell) pairs evaluated: the default, exact sum     */
#define KDEJS_ALGO_DENSE 1    /* fp32 terms / fp64 accumulation brute force over all       */
                              /* N x G pairs (the north-star kernel; FP32-pipe roofline)   */
#define KDEJS_ALGO_DENSE64 2  /* fp64 brute force (cross-check; gaussian parity)           */
#define KDEJS_ALGO_BINNED32 3 /* the culled sum with fp32 terms on the FP32 pipe (sorted     */
                              /* points as 32-bit fixed point, fp64 accumulation across      */
                              /* lanes): exact zeros stay exact, non-zero node densities     */
                              /* agree with BINNED to ~1e-7 relative; nodes it cannot        */
                              /* certify are redone in fp64.  Geometries its fixed-point     */
                              /* format does not cover fall back to BINNED.                  */

/* ---- life cycle -------------------------------------------------------------------- */
int kdejs_version(void);
int kdejs_device_count(int *count);
/* Create (idempotent) the context for `device`: streams, scratch.  Called lazily by every
 * compute entry point; exposed so callers can pay the ~100 ms context cost up front. */
int kdejs_init(int device);
/* NUMA node the GPU's PCIe root hangs off (sysfs), -1 when the platform does not say.  One-process-per-GPU
 * callers bind themselves to that node BEFORE allocating pinned staging memory, so that eight ranks feeding
 * eight GPUs do not cross the socket interconnect (the reference's fork pools have no such concern). */
int kdejs_device_numa_node(int device, int *node);
void kdejs_shutdown(void);
const char *kdejs_last_error(void);

/* ---- the hot path ------------------------------------------------------------------ */
/* KDE_JS_divergence(df1, df2, gamma, eps, log)            KDE_distance.py:104-130
 * (callers: poppunk_mod.py:373, scripts/pairwise_2D_distance.py:54,
 *  scripts/distance_to_gridsearch.py:50, KDE_distance.py:143).
 * a/b: HOST pointers.  resolution/bandwidth/kernel are the reference's hard-coded
 * 100 / 0.03 / epanechnikov (KDE_distance.py:89, :50-51) made explicit.
 * Optional outputs (NULL to skip), each (R*R) doubles on the host:
 *   out_z1/out_z2  the normalised vectors scale_KDE returns    KDE_distance.py:80-95
 *   out_d1/out_d2  raw densities exp(score_samples)            KDE_distance.py:58-64 */
int kdejs_discrepancy(int device, const double *a, int64_t na, const double *b, int64_t nb,
                      int resolution, double bandwidth, double gamma, double eps, int log_flag,
                      int kernel, int algo, double *out_js,
                      double *out_z1, double *out_z2, double *out_d1, double *out_d2);

/* One target against B simulated sets: B independent KDE_JS_divergence(obs, sims[i], ...)
 * calls (the per-simulation call at poppunk_mod.py:373 over a BOLFI batch; the
 * Pool.imap loops at scripts/distance_to_gridsearch.py:109-110).  HOST pointers; sims
 * in pinned memory (kdejs_host_alloc) are copied asynchronously and overlap compute;
 * pageable sims (ordinary numpy arrays) are staged through pinned slots by the library's own
 * threads (KDEJS_STAGE_THREADS, default 4) and reach ~85 % of the pinned rate.
 * out_js: B doubles.  If some sets contain infinity their out_js is NaN and the call
 * returns KDEJS_ERR_INF after finishing the others. */
int kdejs_discrepancy_batch(int device, const double *obs, int64_t n_obs,
                            const double *const *sims, const int64_t *n_sims, int B,
                            int resolution, double bandwidth, double gamma, double eps,
                            int log_flag, int kernel, int algo, double *out_js);

/* process_result's summary vector for a batch of simulations, and the two ELFI nodes built on it, fused behind
 * the discrepancies (poppunk_mod.py:351-382 with calculate_gene_freqs :285-291; Distance("euclidean") and
 * Operation(np.log) :571-572 against the observed vector [0, fc, fi, fr] of :494).  obs/sims as in
 * kdejs_discrepancy_batch; freqs[i] = the simulation's gene-frequency vector (n_freqs[i] doubles, HOST).
 * out_summary: B x 4 = [JS, freq_core, freq_intermediate, freq_rare]; out_d / out_log_d (nullable): B each.
 * The frequencies travel in one upload, one kernel runs behind the last wave, results come back in one copy. */
int kdejs_summary_batch(int device, const double *obs, int64_t n_obs, const double *const *sims,
                        const int64_t *n_sims, const double *const *freqs, const int64_t *n_freqs, int B,
                        double core_threshold, double rare_threshold, const double observed[4], int resolution,
                        double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                        double *out_summary, double *out_d, double *out_log_d);

/* The same summary step for discrepancies that were computed by another entry point -- kdejs_discrepancy_batch_text,
 * when the simulator's tables are parsed on the device (process_result reads them with np.loadtxt,
 * poppunk_mod.py:360-361): js = B discrepancies (HOST); the same kernel as kdejs_summary_batch runs on them, so the
 * vectors equal that call's bit for bit (calculate_gene_freqs :285-291, the result vector :379, the ELFI nodes
 * :571-572). */
int kdejs_summary_from_js(int device, const double *js, const double *const *freqs, const int64_t *n_freqs, int B,
                          double core_threshold, double rare_threshold, const double observed[4],
                          double *out_summary, double *out_d, double *out_log_d);

/* General pairs: out_js[i] = KDE_JS_divergence(sets[ia[i]], sets[ib[i]]) over a pool of
 * n_sets HOST point sets, each uploaded once (the itertools.combinations loop at
 * scripts/pairwise_2D_distance.py:112-118). */
int kdejs_discrepancy_pairs(int device, const double *const *sets, const int64_t *n_pts,
                            int n_sets, const int32_t *ia, const int32_t *ib, int64_t n_pairs,
                            int resolution, double bandwidth, double gamma, double eps,
                            int log_flag, int kernel, int algo, double *out_js);

/* Device-resident variant of kdejs_discrepancy_batch: obs_dev and every sims_dev[i] are
 * DEVICE pointers (16-byte aligned) on `device`, out_js_dev is B doubles on the device,
 * status_dev (nullable) B int32.  All work is enqueued on `stream` (a cudaStream_t, NULL =
 * the library's own stream) and the call returns without synchronising. */
int kdejs_discrepancy_batch_dev(int device, const double *obs_dev, int64_t n_obs,
                                const double *const *sims_dev, const int64_t *n_sims, int B,
                                int resolution, double bandwidth, double gamma, double eps,
                                int log_flag, int kernel, int algo,
                                double *out_js_dev, int32_t *status_dev, void *stream);

/* exp(KernelDensity(bandwidth, kernel).fit(pts).score_samples(get_grid(gmin,gmax,R)[2]))
 * for ALREADY SCALED points -- get_kde + generate_samples, KDE_distance.py:48-64, and the
 * contour density of plot_scatter, poppunk_mod.py:98-105.  HOST pointers. */
int kdejs_kde_density(int device, const double *pts, int64_t n, double gmin, double gmax,
                      int resolution, double bandwidth, int kernel, int algo, double *out_dens);

/* ---- one huge evaluation, grid rows sharded over ranks ------
 * Every rank holds both point sets (HOST pointers) and owns node rows [row_begin,row_end)
 * of the R x R grid (rows = index of column 0).  Phase 1 fits the joint scaler, evaluates
 * the rank's rows and returns its partial normalisers sum(z1**gamma+eps), sum(z2**gamma+eps).
 * The caller all-reduces the two sums (2 doubles) and calls phase 2, which returns the
 * rank's partial relative-entropy sums; JS = sqrt(max(0, (sum of all out_lr[0] + out_lr[1]) / 2)).  (The library
 * accumulates rel_entr(p,m)+rel_entr(q,m) per node -- non-negative, well conditioned -- into out_lr[0]; [1] is 0.) */
int kdejs_shard_phase1(int device, const double *a, int64_t na, const double *b, int64_t nb,
                       int resolution, double bandwidth, double gamma, double eps, int log_flag,
                       int kernel, int algo, int row_begin, int row_end, double out_norm[2]);
/* Same as phase 1 with both point sets already in DEVICE memory (e.g. assembled by an NCCL all-gather of
 * per-rank slices).  Work from other streams that produced them must have completed. */
int kdejs_shard_phase1_dev(int device, const double *a_dev, int64_t na, const double *b_dev, int64_t nb,
                           int resolution, double bandwidth, double gamma, double eps, int log_flag,
                           int kernel, int algo, int row_begin, int row_end, double out_norm[2]);
/* Phase 1 with the row bands chosen by the library: after sorting, the estimated density work of every
 * tile row is summed and the rows are cut into `world` equally heavy bands (real distance clouds are
 * concentrated, equal row counts are not equal work); every rank derives the same cuts from the same
 * data and evaluates band `rank`.  on_device != 0: a/b are DEVICE pointers.  out_rows: the node rows
 * [begin, end) this rank evaluated (may be empty). */
int kdejs_shard_phase1_balanced(int device, const double *a, int64_t na, const double *b, int64_t nb,
                                int on_device, int resolution, double bandwidth, double gamma, double eps,
                                int log_flag, int kernel, int algo, int rank, int world,
                                double out_norm[2], int out_rows[2]);
int kdejs_shard_phase2(int device, const double total_norm[2], double out_lr[2]);

/* ---- one huge evaluation with the POINTS sharded too (band mode) ----------------
 * Replaces the single serial node loop of KDE_distance.py:58-64 over a table no rank holds completely.  Every rank
 * holds a contiguous slice of each table ON ITS DEVICE (rank order = row order) and ends up owning a band of tile
 * columns (4 nodes of grid column 1 each) chosen so that the bands carry equal estimated work.  Only what a band
 * needs crosses NVLink: the points of the cell rows within h of it.  All entry points enqueue on `stream` (NULL = the
 * library's stream); only kdejs_band_plan synchronises.  The caller (poppunk-mod_b200/distributed.py) runs the three
 * small all-gathers and the one all-to-all between the steps:
 *
 *   kdejs_band_minmax_dev   extrema of the own slices                          -> all-gather (10 doubles per rank)
 *   kdejs_band_sort_dev     joint scaler from all extrema (KDE_distance.py:81-86), own slices scaled and counting-sorted
 *                           by cell; out_sorted_dev = (na + nb, 2) scaled points, set 1 after set 0;
 *                           out_block_dev = the rank's count block                 -> all-gather (block_u32 per rank)
 *   kdejs_band_plan         (host sync) work per tile column from the global counts -> kdejs_band_cuts,
 *                           every rank's row starts (sizes of the exchange), moment-cell decision
 *   kdejs_band_cuts / kdejs_band_cell_rows   (pure host) equal-work bands; cell rows a band needs
 *                           -> all-to-all: rank s sends rank d its sorted points of d's cell rows, per set
 *   kdejs_band_phase1_dev   received slices (source after source) merged per cell in rank order = original order,
 *                           density of the band's tiles, partial normalisers       -> all-gather (2 doubles)
 *   kdejs_band_phase2_dev   normalisers folded in rank order, rel_entr partials of the band -> all-gather (3 doubles);
 *                           JS = sqrt(max(0, sum of out_lr[0] + out_lr[1]) / 2); out_lr[2] != 0: an input held infinity.
 * A rank's count block (u32): cellstart[2][C*C + 1] of its sorted slices, then rowstart[2][C + 1], then 16 words
 * for the CUDA IPC handle of the buffer the rank's sorted points live in (zero: none).  kdejs_band_plan returns, per
 * rank, the row starts AND those 16 words (out_rowstart: world x (2 (C + 1) + 16) u32).
 *
 * Exchange over peer memory (ranks of ONE node with NVLink/P2P access; replaces the all-to-all AND the staging copy):
 *   kdejs_band_peer_buffer    this rank's exchange buffer (grown on demand) and its IPC handle: pass the pointer to
 *                             kdejs_band_sort_dev as out_sorted_dev, put the handle words at the end of the count block
 *   kdejs_band_peer_open      after kdejs_band_plan: map every peer's buffer from the gathered handles (cached; only
 *                             a handle that changed is opened again)
 *   kdejs_band_phase1_peer_dev  as kdejs_band_phase1_dev, but the merge kernel reads each source rank's points
 *                             straight from that rank's buffer over NVLink -- the copy is the exchange.
 * Ordering (the caller keeps everything on one stream per rank): a peer's sort precedes its contribution to the
 * all-gather of the count blocks, which precedes this rank's merge; this rank's merge precedes its contribution to
 * the all-gather of the normalisers, which precedes any peer's next sort. */
int kdejs_band_geometry(int resolution, double bandwidth, int kernel, int out[4]);   /* C, NT, block_u32, moments eligible */
int kdejs_band_minmax_dev(int device, const double *a_dev, int64_t na, const double *b_dev, int64_t nb,
                          int log_flag, double eps, double *out_stats_dev, void *stream);
int kdejs_band_sort_dev(int device, const double *a_dev, int64_t na, const double *b_dev, int64_t nb, int resolution,
                        double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                        const double *stats_all_dev, int world, double *out_sorted_dev, uint32_t *out_block_dev,
                        void *stream);
int kdejs_band_plan(int device, const uint32_t *blocks_all_dev, int world, int resolution, double bandwidth,
                    double gamma, double eps, int log_flag, int kernel, int algo, double *out_colwork,
                    uint32_t *out_rowstart, int *out_use_moments, void *stream);
int kdejs_band_cuts(const double *colwork, int nt, int world, int *cuts);
int kdejs_band_cell_rows(int resolution, double bandwidth, int ty_begin, int ty_end, int out_cy[2]);
/* kdejs_band_cuts and every rank's kdejs_band_cell_rows in one call: out_cuts[world + 1], out_rows[world][2] */
int kdejs_band_layout(const double *colwork, int nt, int world, int resolution, double bandwidth, int *out_cuts,
                      int *out_rows);
int kdejs_band_phase1_dev(int device, const double *stage_a_dev, int64_t na, const double *stage_b_dev, int64_t nb,
                          int64_t na_total, int64_t nb_total, const uint32_t *blocks_all_dev, int world,
                          const double *stats_all_dev, int resolution, double bandwidth, double gamma, double eps,
                          int log_flag, int kernel, int algo, int ty_begin, int ty_end, int cy_lo, int cy_hi,
                          int use_moments, double *out_norm_dev, void *stream);
int kdejs_band_phase2_dev(int device, const double *norms_all_dev, int world, double *out_lr_dev, void *stream);
int kdejs_band_peer_buffer(int device, int64_t points, void **out_ptr, uint32_t *out_handle_words /* [16] */);
int kdejs_band_peer_open(int device, const uint32_t *handle_words /* [world][16], host */, int world, int rank);
int kdejs_band_phase1_peer_dev(int device, int64_t na, int64_t nb, int64_t na_total, int64_t nb_total,
                               const uint32_t *blocks_all_dev, int world, const double *stats_all_dev, int resolution,
                               double bandwidth, double gamma, double eps, int log_flag, int kernel, int algo,
                               int ty_begin, int ty_end, int cy_lo, int cy_hi, int use_moments, double *out_norm_dev,
                               void *stream);

/* ---- host memory --------------------------------------------------------------------- */
void *kdejs_host_alloc(size_t bytes);   /* pinned (page-locked) host memory, NULL on failure */
void kdejs_host_free(void *p);

/* ---- input tables -------------------------------------------------------------------- */
/* Multi-threaded ingest of a tab-separated distance table into (rows,2) float64: the LAST TWO
 * columns of every data row (2-column Pansim output, 4-column PopPUNK tables).  Replaces
 * np.loadtxt(file, delimiter='\t') (poppunk_mod.py:360, :264; KDE_distance.py:138-139;
 * scripts/distance_to_gridsearch.py:49) and pandas read_csv + last two columns
 * (scripts/pairwise_2D_distance.py:38-45, poppunk_mod.py:266-267) for this path's inputs.
 * skip_rows: data rows to drop from the top, -1 = drop one row only if it is not numeric.
 * pinned != 0: memory comes from kdejs_host_alloc (async H2D).  Free with kdejs_tsv_free.
 * Errors are reported through kdejs_io_last_error(). */
int kdejs_tsv_load(const char *path, int skip_rows, int pinned, int threads, double **out_pts,
                   int64_t *out_rows);
/* The same table parsed into the CALLER's buffer of capacity_rows x 2 doubles -- e.g. one slot of a page-locked
 * (B,N,2) arena that a batch call then uploads as a whole (a sweep over thousands of files: no allocation per
 * file).  *out_rows = data rows of the file, also when they do not fit: then nothing is written and the call
 * returns KDEJS_ERR_CAPACITY. */
int kdejs_tsv_load_into(const char *path, int skip_rows, int threads, double *dst, int64_t capacity_rows,
                        int64_t *out_rows);
void kdejs_tsv_free(double *pts, int pinned);
/* Every number of a small text file as one vector (the simulator's gene frequencies, poppunk_mod.py:361): separated by
 * blanks, tabs or newlines; '#' comments; pandas' missing-value spellings are NaN.  Free with kdejs_tsv_free(p, 0). */
int kdejs_numbers_load(const char *path, double **out, int64_t *out_n);
/* First byte of the data rows of a table that is already in memory (same header rule as kdejs_tsv_load). */
int64_t kdejs_tsv_data_offset(const char *text, int64_t bytes, int skip_rows);
/* n files read into slots of slot_bytes bytes (dst + i * slot_bytes), `threads` at a time; out_bytes[i] = bytes read,
 * -(file size) when the file does not fit its slot, -1 when it cannot be opened.  No parsing. */
int kdejs_files_read(const char *const *paths, int n, char *dst, int64_t slot_bytes, int64_t *out_bytes, int threads);

/* ---- sweeps over tables on disk: the TEXT goes to the device ------------------------- */
/* [KDE_JS_divergence(obs, table_i) for i in range(B)] where table_i is given as the bytes of its file
 * (texts[i], text_bytes[i]; page-locked memory for asynchronous copies).  The text is uploaded as it is and parsed on
 * the device (last two tab-separated columns, skip_rows as kdejs_tsv_load; csrc/kdejs_tsv.cuh) -- the loop of
 * scripts/distance_to_gridsearch.py:48-55, :109-110 without a host parser in it.  Numbers get the bits float() gives
 * them (csrc/kdejs_decimal.h).  out_rows[i] = data rows of table i; out_bounds[i] = (min, max) of column 0 and of
 * column 1 as np.min / np.max give them (the bounds test of distance_to_gridsearch.py:112-125).  out_status[i]:
 *   KDEJS_TABLE_OK    out_js[i] is the discrepancy
 *   KDEJS_TABLE_HOST  the table holds something the device parser does not cover (a missing-value spelling, inf / nan,
 *                     more than 19 digits, a malformed row) or no data rows: out_js[i] is NaN -- read that file with
 *                     kdejs_tsv_load (which reports malformed rows) and evaluate it through the host entry points
 *   KDEJS_TABLE_INF   the table parsed, but holds an infinity the path rejects (KDEJS_ERR_INF for a host call) */
#define KDEJS_TABLE_OK 0
#define KDEJS_TABLE_HOST 1
#define KDEJS_TABLE_INF 2
/* One table parsed by the device parser, rows copied back to out_pts (capacity_rows x 2 doubles): what
 * kdejs_tsv_load_into gives, computed on the GPU (parity tests; callers that want the rows of a table they hold as
 * text).  *out_status = KDEJS_TABLE_OK or KDEJS_TABLE_HOST (then the rows are not to be used). */
int kdejs_tsv_parse_text(int device, const char *text, int64_t bytes, int skip_rows, double *out_pts,
                         int64_t capacity_rows, int64_t *out_rows, int32_t *out_status);
int kdejs_discrepancy_batch_text(int device, const double *obs, int64_t n_obs, const char *const *texts,
                                 const int64_t *text_bytes, int B, int skip_rows, int resolution, double bandwidth,
                                 double gamma, double eps, int log_flag, int kernel, int algo, double *out_js,
                                 int64_t *out_rows, int32_t *out_status, double *out_bounds /* nullable, [B][4] */);
const char *kdejs_io_last_error(void);

/* ---- measurement support (bench.py, tests) ------------------------------------------- */
/* Per-kernel CUDA-event timing on the launching stream.  which: 0 minmax, 1 scale+count,
 * 2 scan, 3 place, 4 density (the dominant kernel), 5 divergence, 6 tile classification,
 * 7 cell moments (fine grids only), 8 fp64 redo pass of KDEJS_ALGO_BINNED32, 9 summary vector,
 * 10 table text: count + scan, 11 table text: parse (+ bounds) (kdejs_discrepancy_batch_text).
 * Reading synchronises. */
#define KDEJS_NUM_KERNELS 12
int kdejs_profile_enable(int device, int on);   /* 0 off, 1 every kernel, 2 the density kernel only */
int kdejs_profile_read(int device, int which, double *ms_total, int64_t *launches);
int kdejs_profile_reset(int device);
/* Kernels launched by this library on `device` since init / last reset. */
int kdejs_launch_count(int device, int64_t *count, int reset);
/* Work counters of the density launches since the last reset: pair_tests = (point, node)
 * kernel evaluations executed; in_support is filled only by the CPU-side estimate (0 here). */
int kdejs_work_counters(int device, double *pair_tests, int reset);
/* KDEJS_ALGO_BINNED32 bookkeeping since the last reset, filled only while profiling is enabled: out[0] tiles,
 * [1] tiles walked a second time (a node's fp32 sum below the certification threshold), [2] / [3] warp trips of
 * the first / second walks, [4] tiles re-queued for the fp64 kernel. */
int kdejs_fp32_stats(int device, double out[5], int reset);
/* The candidate loops of the two culled density kernels over an L1-resident run, no per-tile work: (point, node)
 * pair tests per second.  which: 0 fp64 loop (k_density_binned), 1 fp32 loop without, 2 with the periodic fold of
 * the fp32 partial sums into fp64 (k_density_binned32).  The ceiling the kernels' candidate work can approach. */
int kdejs_loop_peak(int device, int which, double *pair_tests_per_s);
/* Issue-rate microbenchmarks used as roofline denominators.  which: 0 FFMA fp32,
 * 1 DFMA fp64, 2 packed FFMA2 (f32x2), 3 / 4 DFMA interleaved with one / two independent ALU-pipe
 * instructions per DFMA (reports the DFMA rate only: does a co-issued instruction slow the fp64 pipe?).
 * Returns achieved TFLOP/s (2 flop per lane-FMA). */
int kdejs_peak_flops(int device, int which, double *tflops, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* KDEJS_H */
