"""One process per GPU (torchrun): sharding of the two data-parallel shapes of the path.

* Batches / sweeps (BOLFI batches, Latin-hypercube grids, all-pairs tables): evaluations are
  independent -- each has its own joint scaler (KDE_distance.py:81-86) -- so items are dealt to
  ranks and only the scalar discrepancies are gathered.  No collective touches the data path.
* One huge evaluation (config C5: 10 M pairs on a 1024^2 grid): node rows are split over ranks;
  the path has two real exchange steps, the normalisers (2 doubles) and the relative-entropy
  partials (2 doubles).  They are all-gathered and summed in rank order, so every rank ends with
  the same bit pattern.

torch.distributed is used for plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes

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
                            group=None, device=None, sliced=False, phase1=None, phase2=None):
    """KDE_JS_divergence of ONE pair of (large) point sets with the grid's node rows split over
    the ranks of `group`.  Returns the same np.float64 on all ranks.

    sliced=False: every rank passes both complete sets (each rank uploads everything).
    sliced=True : every rank passes only ITS contiguous slice of each set (rank order = row order of the
                  table); slices are exchanged GPU-to-GPU with one all-gather per set."""
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = kd.default_device() if device is None else int(device)
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
