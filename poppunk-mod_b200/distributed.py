"""One process per GPU (torchrun): sharding of the two data-parallel shapes of the path.

* Batches / sweeps (BOLFI batches, Latin-hypercube grids, all-pairs tables): evaluations are
  independent -- each has its own joint scaler (KDE_distance.py:81-86) -- so items are dealt to
  ranks and only the scalar discrepancies are gathered.  No collective touches the data path.
* One huge evaluation (config C5: 10 M pairs on a 1024^2 grid): the grid is split into bands of tile
  columns, one per rank, AND the points are sharded: every rank holds a slice of each table, sorts
  it locally, and receives only the points of the cell rows within h of its band
  (band_sharded_discrepancy).  On the GPUs of one node the band's merge kernel reads those rows straight
  out of the peers' sorted arrays over NVLink (CUDA IPC mappings, kdejs_band_peer_*): the copy that puts
  them in order IS the exchange.  Otherwise one grouped NCCL all-to-all of contiguous slices.  The remaining
  exchange steps are tiny: the extrema (10 doubles), the per-cell counts, the normalisers (2
  doubles) and the relative-entropy partials (3 doubles); each is all-gathered and folded in rank
  order, so every rank ends with the same bit pattern.  (row_sharded_discrepancy with complete
  tables on every rank is the older replicate-and-sort variant.)

torch.distributed is used for plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _native as nat
from . import KDE_distance as kd

TILE = 4   # kdejs_shard_phase1 wants row ranges aligned to the kernel's 4x4-node tiles


def shard_items(n_items: int, rank: int, world: int):
    """Indices of the evaluations rank `rank` owns (round-robin keeps ragged batches balanced)."""
    return list(range(rank, n_items, world))


def row_ranges(resolution: int, world: int):
    """Contiguous node-row ranges, one per rank, each starting on a multiple of 4."""
    tiles = (resolution + TILE - 1) // TILE
    cuts = [min(resolution, TILE * ((tiles * r) // world)) for r in range(world)] + [resolution]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _dist():
    import torch.distributed as dist

    return dist


def _all_gather_f64(values, group=None, device=None):
    """all_gather of a small float64 vector; returns (world, len) as numpy, rank-ordered."""
    import torch

    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(values, dtype=np.float64)[None, :]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", device if device is not None else kd.default_device()) if backend == "nccl" \
        else torch.device("cpu")
    t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy() for o in out])


def batch_discrepancy_sharded(obs, sims, gamma=1.0, eps=1e-12, log=False, *, group=None, compute=None, **kw):
    """Every rank passes the SAME list `sims` (or a list whose foreign entries are None); rank r
    evaluates items r, r+world, ... on its own GPU and all ranks return the full (B,) result."""
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = len(sims)
    mine = shard_items(B, rank, world)
    compute = compute or kd.KDE_JS_divergence_batch
    local = np.asarray(compute(obs, [sims[i] for i in mine], gamma, eps, log, **kw), dtype=np.float64)
    width = (B + world - 1) // world
    padded = np.full(width, np.nan)
    padded[:len(local)] = local
    gathered = _all_gather_f64(padded, group)
    out = np.empty(B, dtype=np.float64)
    for r in range(gathered.shape[0]):
        idx = shard_items(B, r, world)
        out[idx] = gathered[r, :len(idx)]
    return out


def _phase1_gpu(df1, df2, common, lo, hi, device):
    a, b = nat.as_points(df1, "df1"), nat.as_points(df2, "df2")
    n = (ctypes.c_double * 2)()
    nat.check(nat.lib().kdejs_shard_phase1(device, nat.ptr(a), a.shape[0], nat.ptr(b), b.shape[0], *common,
                                           lo, hi, n))
    return np.array([n[0], n[1]])


def _phase2_gpu(totals, device):
    t = (ctypes.c_double * 2)(float(totals[0]), float(totals[1]))
    lr = (ctypes.c_double * 2)()
    nat.check(nat.lib().kdejs_shard_phase2(device, t, lr))
    return np.array([lr[0], lr[1]])


def _gather_slices(local, group, device):
    """All ranks contribute a (n_r, 2) slice; every rank gets the rank-ordered concatenation.  NCCL:
    the slice goes H2D once (1/world of the table per rank), the exchange runs GPU-to-GPU over
    NVLink/NVSwitch, and the result is returned as a device tensor.  gloo: numpy (CPU tests)."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    loc = nat.as_points(local, "slice")
    sizes = _all_gather_f64(np.array([loc.shape[0]], dtype=np.float64), group, device)[:, 0].astype(np.int64)
    nmax = int(sizes.max())
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", device)
        pad = torch.zeros((nmax, 2), dtype=torch.float64, device=dev)
        pad[:loc.shape[0]].copy_(torch.from_numpy(loc), non_blocking=True)
        out = torch.empty((world, nmax, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, pad, group=group)
        full = out.reshape(-1, 2) if bool((sizes == nmax).all()) else \
            torch.cat([out[r, :int(sizes[r])] for r in range(world)], dim=0)
        torch.cuda.synchronize(dev)
        return full.contiguous()
    pad = torch.zeros((nmax, 2), dtype=torch.float64)
    pad[:loc.shape[0]] = torch.from_numpy(loc)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([out[r][:int(sizes[r])].numpy() for r in range(world)], axis=0)


def row_sharded_discrepancy(df1, df2, gamma=1.0, eps=1e-12, log=False, *, resolution=100, bandwidth=0.03,
                            group=None, device=None, sliced=False, phase1=None, phase2=None, algo=None,
                            gather_all=False):
    """KDE_JS_divergence of ONE pair of (large) point sets with the grid's node rows split over
    the ranks of `group`.  Returns the same np.float64 on all ranks.

    sliced=False: every rank passes both complete sets (each rank uploads everything).
    sliced=True : every rank passes only ITS contiguous slice of each set (rank order = row order of the
                  table).  On GPUs the points stay sharded (band_sharded_discrepancy: local sort, only the rows a
                  band needs are exchanged); gather_all=True forces the older variant that all-gathers both
                  tables to every rank (also the path of the CPU tests, which inject phase1/phase2)."""
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = kd.default_device() if device is None else int(device)
    if sliced and phase1 is None and (world == 1 or dist.get_backend(group) == "nccl") and not gather_all:
        # the points stay sharded: local sort, band-wise point exchange (band mode)
        return band_sharded_discrepancy(df1, df2, gamma, eps, log, resolution=resolution, bandwidth=bandwidth,
                                        group=group, device=dev, algo=algo)
    common = kd._common(resolution, bandwidth, gamma, eps, log, "epanechnikov", "binned")
    if sliced and dist.is_initialized():
        df1, df2 = _gather_slices(df1, group, dev), _gather_slices(df2, group, dev)
    if phase1 is not None:                       # injected compute (CPU tests): equal row counts
        lo, hi = row_ranges(resolution, world)[rank]
        norms = phase1(df1, df2, lo, hi) if lo < hi else np.zeros(2)
    else:
        # the library cuts the rows into bands of equal estimated work (same cuts on every rank)
        n = (ctypes.c_double * 2)()
        rows = (ctypes.c_int * 2)()
        if isinstance(df1, np.ndarray) or not hasattr(df1, "data_ptr"):
            a, b = nat.as_points(df1, "df1"), nat.as_points(df2, "df2")
            pa, pb, on_dev = nat.ptr(a), nat.ptr(b), 0
            na, nb = a.shape[0], b.shape[0]
        else:                                    # device tensors from the all-gather
            pa, pb, on_dev = df1.data_ptr(), df2.data_ptr(), 1
            na, nb = df1.shape[0], df2.shape[0]
        nat.check(nat.lib().kdejs_shard_phase1_balanced(dev, pa, na, pb, nb, on_dev, *common, rank, world, n, rows))
        norms = np.array([n[0], n[1]])
        lo, hi = rows[0], rows[1]
    totals = _all_gather_f64(norms, group, dev).sum(axis=0)          # rank order: deterministic
    if lo < hi:
        lr = phase2(totals) if phase2 else _phase2_gpu(totals, dev)
    else:
        lr = np.zeros(2)
    parts = _all_gather_f64(lr, group, dev)
    js2 = (parts[:, 0].sum() + parts[:, 1].sum()) / 2.0
    if js2 < 0.0:            # rounding noise around an exact 0; NaN compares false and stays NaN
        js2 = 0.0
    return np.float64(np.sqrt(js2))


# ---------------------------------------------------------------------------------------------------
# Band mode: the points are sharded too (C ABI: kdejs_band_*, include/kdejs.h)

class _GpuBandOps:
    """The device steps of band mode on this rank's GPU.  Everything -- the library's kernels, torch's own copies and
    the NCCL collectives torch issues in between -- runs on ONE dedicated torch stream (`run()` makes it current), so
    the steps are ordered without any host synchronisation.  It must be a real stream: torch's default stream has
    the NULL handle, which the C ABI reads as "the library's own stream", and that one is not ordered with torch."""

    def __init__(self, device, common):
        import torch

        self.torch, self.dev, self.common = torch, int(device), common
        self.L = nat.lib()
        geo = (ctypes.c_int * 4)()
        nat.check(self.L.kdejs_band_geometry(common[0], common[1], common[5], geo))
        self.C, self.NT, self.block_u32 = geo[0], geo[1], geo[2]
        self.tdev = torch.device("cuda", self.dev)
        self.stream = _band_stream(torch, self.tdev)
        # exchange over peer memory when every rank can: one node (torchrun exports LOCAL_WORLD_SIZE; all ranks read
        # the same value, so all decide alike), at most MAX_PEERS ranks, not switched off
        self.peer = os.environ.get("KDEJS_BAND_PEER", "1") != "0" and \
            os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) == os.environ.get("WORLD_SIZE", "1")
        self.handles = np.zeros((1, HANDLE_WORDS), dtype=np.uint32)

    def run(self):
        """Context manager: inputs produced on the caller's current stream are waited for, then the band stream is
        made current for everything inside."""
        self.stream.wait_stream(self.torch.cuda.current_stream(self.tdev))
        return self.torch.cuda.stream(self.stream)

    def _stream(self):
        return self.stream.cuda_stream

    def to_device(self, x, name):
        torch = self.torch
        if hasattr(x, "data_ptr"):
            if x.dtype != torch.float64 or x.dim() != 2 or x.shape[1] != 2 or not x.is_contiguous() or not x.is_cuda:
                raise ValueError(f"{name}: device slices must be contiguous (n,2) float64 CUDA tensors")
            return x
        return torch.from_numpy(nat.as_points(x, name)).to(self.tdev, non_blocking=True)

    def empty(self, shape, dtype="f8"):
        t = self.torch
        return t.empty(shape, dtype={"f8": t.float64, "i4": t.int32}[dtype], device=self.tdev)

    def minmax(self, a, b):
        out = self.empty(10)
        nat.check(self.L.kdejs_band_minmax_dev(self.dev, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0],
                                               self.common[4], self.common[3], out.data_ptr(), self._stream()))
        return out

    def _peer_buffer(self, points):
        """This rank's exchange buffer as a tensor view, and the device copy of its IPC handle words."""
        torch = self.torch
        ptr, words = ctypes.c_void_p(), np.zeros(HANDLE_WORDS, dtype=np.uint32)
        nat.check(self.L.kdejs_band_peer_buffer(self.dev, int(points), ctypes.byref(ptr), words.ctypes.data))
        key = words.tobytes()
        cached = _HANDLE_DEV.get(self.dev)
        if cached is None or cached[0] != key:                      # uploaded once per (re)allocation of the buffer
            cached = _HANDLE_DEV[self.dev] = (key, torch.from_numpy(words.view(np.int32)).to(self.tdev))
        view = torch.as_tensor(_DevView(ptr.value, (max(int(points), 1), 2)), device=self.tdev)
        return view[:int(points)], cached[1]

    def sort(self, a, b, stats_all, world):
        block = self.empty(self.block_u32, "i4")
        use_peer = self.peer and 1 < world <= MAX_PEERS and not _PEER_BROKEN.get(self.dev)
        if use_peer:
            try:
                srt, handle = self._peer_buffer(a.shape[0] + b.shape[0])
                block[-HANDLE_WORDS:] = handle
            except RuntimeError:                 # no IPC export here: the zero handle makes every rank use NCCL
                self.peer = use_peer = False
        if not use_peer:
            srt = self.empty((a.shape[0] + b.shape[0], 2))
            block[-HANDLE_WORDS:] = 0
        nat.check(self.L.kdejs_band_sort_dev(self.dev, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], *self.common,
                                             stats_all.data_ptr(), world, srt.data_ptr(), block.data_ptr(),
                                             self._stream()))
        return srt, block

    def plan(self, blocks_all, world):
        colwork = np.empty(self.NT, dtype=np.float64)
        out = np.empty((world, 2 * (self.C + 1) + HANDLE_WORDS), dtype=np.uint32)
        use_m = ctypes.c_int(0)
        nat.check(self.L.kdejs_band_plan(self.dev, blocks_all.data_ptr(), world, *self.common, colwork.ctypes.data,
                                         out.ctypes.data, ctypes.byref(use_m), self._stream()))
        self.handles = np.ascontiguousarray(out[:, -HANDLE_WORDS:])             # every rank's IPC handle words
        rowstart = out[:, :-HANDLE_WORDS].reshape(world, 2, self.C + 1)
        return colwork, rowstart.astype(np.int64), use_m.value

    def peer_ready(self, rank, world, group=None):
        """True when every rank exported its sorted points and every rank could map them.  All ranks see the same
        handles, so all take the same branch; when the set of handles is new (first call, or a buffer had to grow)
        the ranks also agree -- one tiny all-gather -- that every mapping succeeded, and fall back to the NCCL
        exchange together, for good, if one did not."""
        if world == 1 or _PEER_BROKEN.get(self.dev) or not self.handles.any(axis=1).all():
            return False
        key = (world, rank, self.handles.tobytes())
        if _PEER_OPEN.get(self.dev) == key:
            return True
        rc = self.L.kdejs_band_peer_open(self.dev, self.handles.ctypes.data, world, rank)
        ok = self.torch.tensor([1 if rc == 0 else 0], dtype=self.torch.int32, device=self.tdev)
        if not bool(_gather_cat(ok, group, world).cpu().numpy().all()):
            _PEER_BROKEN[self.dev] = True
            return False
        _PEER_OPEN[self.dev] = key
        return True

    def phase1_peer(self, na, nb, na_tot, nb_tot, blocks_all, stats_all, world, ty, cy, use_m):
        out = self.empty(2)
        nat.check(self.L.kdejs_band_phase1_peer_dev(
            self.dev, int(na), int(nb), int(na_tot), int(nb_tot), blocks_all.data_ptr(), world, stats_all.data_ptr(),
            *self.common, ty[0], ty[1], cy[0], cy[1], use_m, out.data_ptr(), self._stream()))
        return out

    def phase1(self, stage_a, stage_b, na_tot, nb_tot, blocks_all, stats_all, world, ty, cy, use_m):
        out = self.empty(2)
        nat.check(self.L.kdejs_band_phase1_dev(
            self.dev, stage_a.data_ptr() if stage_a.shape[0] else None, stage_a.shape[0],
            stage_b.data_ptr() if stage_b.shape[0] else None, stage_b.shape[0], int(na_tot), int(nb_tot),
            blocks_all.data_ptr(), world, stats_all.data_ptr(), *self.common, ty[0], ty[1], cy[0], cy[1], use_m,
            out.data_ptr(), self._stream()))
        return out

    def phase2(self, norms_all, world):
        out = self.empty(3)
        nat.check(self.L.kdejs_band_phase2_dev(self.dev, norms_all.data_ptr(), world, out.data_ptr(), self._stream()))
        return out

    def to_host(self, t):
        return t.cpu().numpy()


HANDLE_WORDS = 16      # a CUDA IPC handle, as u32 words at the end of a rank's count block (include/kdejs.h)
MAX_PEERS = 16         # ranks the merge kernel can pull from (kMaxPeers in csrc/kdejs_kernels.cuh)


class _DevView:
    """A device allocation of the library as something torch.as_tensor can wrap without a copy."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


_BAND_STREAMS = {}
_HANDLE_DEV = {}
_PEER_OPEN = {}        # device -> (world, rank, handles) of the mappings that are open
_PEER_BROKEN = {}      # device -> True once a rank could not map a peer: NCCL exchange from then on


def _band_stream(torch, tdev):
    """One long-lived non-default stream per device for band mode (creating a stream per call costs ~10 us)."""
    key = (tdev.type, tdev.index)
    if key not in _BAND_STREAMS:
        _BAND_STREAMS[key] = torch.cuda.Stream(tdev)
    return _BAND_STREAMS[key]


def band_cuts(colwork, world):
    """Equal-work bands of tile columns: cuts[r] .. cuts[r+1] for rank r (kdejs_band_cuts, pure host)."""
    cw = np.ascontiguousarray(colwork, dtype=np.float64)
    cuts = (ctypes.c_int * (world + 1))()
    nat.check(nat.lib().kdejs_band_cuts(cw.ctypes.data, len(cw), world, cuts))
    return list(cuts)


def band_cell_rows(resolution, bandwidth, ty_begin, ty_end):
    """Cell rows [lo, hi] whose points can reach tile columns [ty_begin, ty_end) (hi < lo: empty band)."""
    cy = (ctypes.c_int * 2)()
    nat.check(nat.lib().kdejs_band_cell_rows(int(resolution), float(bandwidth), int(ty_begin), int(ty_end), cy))
    return cy[0], cy[1]


def band_layout(colwork, world, resolution, bandwidth):
    """band_cuts and every rank's band_cell_rows in one call: (cuts, rows) as plain lists."""
    cw = np.ascontiguousarray(colwork, dtype=np.float64)
    cuts = np.empty(world + 1, dtype=np.int32)
    rows = np.empty((world, 2), dtype=np.int32)
    nat.check(nat.lib().kdejs_band_layout(cw.ctypes.data, len(cw), world, int(resolution), float(bandwidth),
                                          cuts.ctypes.data, rows.ctypes.data))
    return cuts.tolist(), [tuple(r) for r in rows.tolist()]


def _gather_cat(t, group, world):
    """Rank-ordered concatenation of equally sized 1-D/2-D tensors from all ranks (no host round trip)."""
    dist = _dist()
    if world == 1:
        return t
    import torch

    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def _exchange(pairs, group, rank, world):
    """pairs: [(send, recv), ...] with send[d] -> rank d and recv[s] <- rank s for every pair (one pair per point
    set).  All transfers of all pairs go out as ONE group of point-to-point operations (one NCCL group launch, not
    one collective per set); slices of a `send` may overlap (halo rows go to two neighbours); empty slices are
    skipped on both sides (both know the count from the gathered row starts)."""
    dist = _dist()
    for send, recv in pairs:
        if recv[rank].shape[0]:
            recv[rank].copy_(send[rank])
    if world == 1:
        return
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    ops = []
    for k in range(1, world):
        d, s_ = (rank + k) % world, (rank - k) % world
        for send, recv in pairs:
            if send[d].shape[0]:
                ops.append(dist.P2POp(dist.isend, send[d].contiguous(), peer(d), group))
            if recv[s_].shape[0]:
                ops.append(dist.P2POp(dist.irecv, recv[s_], peer(s_), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def band_sharded_discrepancy(a_slice, b_slice, gamma=1.0, eps=1e-12, log=False, *, resolution=100, bandwidth=0.03,
                             algo=None, group=None, device=None, ops=None, info=None):
    """KDE_JS_divergence of ONE pair of (large) tables whose rows are sharded over the ranks of `group`: every rank
    passes ITS contiguous slice of each table (rank order = row order; numpy, pinned numpy or a CUDA tensor).
    Returns the same np.float64 on all ranks; raises ValueError on infinity like the single-GPU call.
    `info` (a dict) receives the band this rank owned and the sizes of the exchange."""
    dist = _dist()
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    dev = kd.default_device() if device is None else int(device)
    common = kd._common(resolution, bandwidth, gamma, eps, log, "epanechnikov", algo)
    if ops is None:
        ops = _GpuBandOps(dev, common)
    with ops.run():
        return _band_steps(ops, a_slice, b_slice, resolution, bandwidth, group, rank, world, info)


def _band_steps(ops, a_slice, b_slice, resolution, bandwidth, group, rank, world, info):
    import time

    timing = info is not None and info.get("timing")           # per-step wall clock with a device sync after each step
    marks = []

    def mark(name):
        if timing:
            ops.torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    a, b = ops.to_device(a_slice, "df1 slice"), ops.to_device(b_slice, "df2 slice")
    na = a.shape[0]
    mark("upload")
    stats_all = _gather_cat(ops.minmax(a, b), group, world)                    # joint scaler: 10 doubles per rank
    mark("minmax+gather")
    srt, block = ops.sort(a, b, stats_all, world)
    mark("local sort")
    blocks_all = _gather_cat(block, group, world)                              # per-cell counts of every rank
    mark("gather counts")
    colwork, rs, use_m = ops.plan(blocks_all, world)                           # the one host synchronisation
    mark("plan")
    # (the GPU idles between the plan's synchronisation and the launches of phase 1: keep this stretch short)
    cuts, rows = band_layout(colwork, world, resolution, bandwidth)
    lo_me, hi_me = rows[rank]
    n_tot = rs[:, :, -1].sum(axis=0)
    recv = (rs[:, :, hi_me + 1] - rs[:, :, lo_me]).T.tolist() if hi_me >= lo_me else [[0] * world, [0] * world]
    over_peer_memory = bool(getattr(ops, "peer_ready", None)) and ops.peer_ready(rank, world, group)
    if over_peer_memory:
        # the merge kernel pulls the rows it needs out of every rank's sorted array over NVLink
        mark("exchange")
        norm = ops.phase1_peer(sum(recv[0]), sum(recv[1]), n_tot[0], n_tot[1], blocks_all, stats_all, world,
                               (cuts[rank], cuts[rank + 1]), (lo_me, hi_me), use_m)
    else:
        stages, pairs = [], []
        for st in (0, 1):
            base = 0 if st == 0 else na
            send = [srt[base + rs[rank, st, lo]: base + rs[rank, st, hi + 1]] if hi >= lo else srt[0:0] for lo, hi in rows]
            stage = ops.empty((sum(recv[st]), 2))
            offs = np.concatenate([[0], np.cumsum(recv[st])])
            pairs.append((send, [stage[offs[s_]:offs[s_ + 1]] for s_ in range(world)]))
            stages.append(stage)
        _exchange(pairs, group, rank, world)
        mark("exchange")
        norm = ops.phase1(stages[0], stages[1], n_tot[0], n_tot[1], blocks_all, stats_all, world,
                          (cuts[rank], cuts[rank + 1]), (lo_me, hi_me), use_m)
    mark("merge+density")
    lr = ops.phase2(_gather_cat(norm, group, world), world)
    parts = ops.to_host(_gather_cat(lr, group, world)).reshape(world, 3)        # final synchronisation
    mark("divergence+gathers")
    if timing:
        info["timing_ms"] = {n: 1e3 * (t - marks[i][1]) for i, (n, t) in enumerate(marks[1:])}
    if info is not None:
        info.update({"tile_columns": (cuts[rank], cuts[rank + 1]), "cell_rows": (lo_me, hi_me),
                     "points_received": (sum(recv[0]), sum(recv[1])), "moment_cells": bool(use_m),
                     "exchange": "peer memory" if over_peer_memory else "nccl"})
    if parts[:, 2].any():
        raise ValueError("Input X contains infinity or a value too large for dtype('float64').")
    js2 = (parts[:, 0].sum() + parts[:, 1].sum()) / 2.0
    if js2 < 0.0:
        js2 = 0.0
    return np.float64(np.sqrt(js2))
