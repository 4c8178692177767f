"""ctypes binding of libkdejs.so (include/kdejs.h).  No torch, no numpy fallbacks: if the CUDA
library is missing or no sm_100 device is usable, every compute call raises."""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkdejs.so")

KDEJS_OK, KDEJS_ERR_CUDA, KDEJS_ERR_ARG, KDEJS_ERR_INF, KDEJS_ERR_CAPACITY = 0, 1, 2, 3, 4
KDEJS_TABLE_OK, KDEJS_TABLE_HOST, KDEJS_TABLE_INF = 0, 1, 2
KERNELS = {"epanechnikov": 0, "gaussian": 1}
ALGOS = {"binned": 0, "dense": 1, "dense64": 2, "binned32": 3}
KERNEL_NAMES = ("minmax", "scale_count", "scan", "place", "density", "divergence", "classify", "moments",
                "redo64", "summary", "tsv_count", "tsv_parse")

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_dpp = ctypes.POINTER(ctypes.c_void_p)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)

_lib = None
_lock = threading.Lock()

_COMMON = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
           ctypes.c_int]  # resolution, bandwidth, gamma, eps, log_flag, kernel, algo

_SIGNATURES = {
    "kdejs_version": (ctypes.c_int, []),
    "kdejs_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "kdejs_device_numa_node": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "kdejs_init": (ctypes.c_int, [ctypes.c_int]),
    "kdejs_shutdown": (None, []),
    "kdejs_last_error": (ctypes.c_char_p, []),
    "kdejs_discrepancy": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_int64] + _COMMON + [_c_dp] + [ctypes.c_void_p] * 4),
    "kdejs_discrepancy_batch": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, _c_dpp, _c_i64p,
                                               ctypes.c_int] + _COMMON + [ctypes.c_void_p]),
    "kdejs_summary_batch": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, _c_dpp, _c_i64p, _c_dpp, _c_i64p,
                                           ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p] + _COMMON +
                            [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "kdejs_summary_from_js": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _c_dpp, _c_i64p, ctypes.c_int, ctypes.c_double,
                                             ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "kdejs_discrepancy_pairs": (ctypes.c_int, [ctypes.c_int, _c_dpp, _c_i64p, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int64] + _COMMON + [ctypes.c_void_p]),
    "kdejs_discrepancy_batch_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, _c_dpp, _c_i64p,
                                                   ctypes.c_int] + _COMMON + [ctypes.c_void_p, ctypes.c_void_p,
                                                                              ctypes.c_void_p]),
    "kdejs_kde_density": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p]),
    "kdejs_shard_phase1": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                          ctypes.c_int64] + _COMMON + [ctypes.c_int, ctypes.c_int, _c_dp]),
    "kdejs_shard_phase1_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                              ctypes.c_int64] + _COMMON + [ctypes.c_int, ctypes.c_int, _c_dp]),
    "kdejs_shard_phase1_balanced": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                   ctypes.c_int64, ctypes.c_int] + _COMMON +
                                    [ctypes.c_int, ctypes.c_int, _c_dp, ctypes.POINTER(ctypes.c_int)]),
    "kdejs_shard_phase2": (ctypes.c_int, [ctypes.c_int, _c_dp, _c_dp]),
    "kdejs_band_geometry": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "kdejs_band_minmax_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "kdejs_band_sort_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_int64] + _COMMON + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                                        ctypes.c_void_p, ctypes.c_void_p]),
    "kdejs_band_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + _COMMON +
                        [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    "kdejs_band_cuts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "kdejs_band_cell_rows": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int)]),
    "kdejs_band_layout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "kdejs_band_phase1_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_int, ctypes.c_void_p] + _COMMON +
                              [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_void_p]),
    "kdejs_band_phase2_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "kdejs_band_peer_buffer": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p),
                                              ctypes.c_void_p]),
    "kdejs_band_peer_open": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "kdejs_band_phase1_peer_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] +
                                   _COMMON + [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    "kdejs_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "kdejs_host_free": (None, [ctypes.c_void_p]),
    "kdejs_tsv_load": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_void_p), _c_i64p]),
    "kdejs_tsv_load_into": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int64, _c_i64p]),
    "kdejs_tsv_free": (None, [ctypes.c_void_p, ctypes.c_int]),
    "kdejs_numbers_load": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), _c_i64p]),
    "kdejs_tsv_data_offset": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "kdejs_files_read": (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                        _c_i64p, ctypes.c_int]),
    "kdejs_tsv_parse_text": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int64, _c_i64p, ctypes.POINTER(ctypes.c_int32)]),
    "kdejs_discrepancy_batch_text": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, _c_dpp, _c_i64p,
                                                    ctypes.c_int, ctypes.c_int] + _COMMON +
                                     [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "kdejs_io_last_error": (ctypes.c_char_p, []),
    "kdejs_profile_enable": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "kdejs_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_dp, _c_i64p]),
    "kdejs_profile_reset": (ctypes.c_int, [ctypes.c_int]),
    "kdejs_launch_count": (ctypes.c_int, [ctypes.c_int, _c_i64p, ctypes.c_int]),
    "kdejs_work_counters": (ctypes.c_int, [ctypes.c_int, _c_dp, ctypes.c_int]),
    "kdejs_peak_flops": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_dp, _c_dp]),
    "kdejs_fp32_stats": (ctypes.c_int, [ctypes.c_int, _c_dp, ctypes.c_int]),
    "kdejs_loop_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_dp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load libkdejs.so (dlopen only -- CUDA is not touched until the first compute call, so this
    is safe to do before a fork)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with poppunk-mod_b200/csrc/build.sh "
                        "(there is no CPU fallback)")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    f = getattr(L, name)
                    f.restype = res
                    f.argtypes = args
                _lib = L
    return _lib


def check(rc: int):
    if rc == KDEJS_OK:
        return
    msg = (lib().kdejs_last_error() or b"").decode("utf-8", "replace")
    if rc in (KDEJS_ERR_ARG, KDEJS_ERR_INF):
        raise ValueError(msg)
    raise RuntimeError(f"kdejs: {msg}")


def device_count() -> int:
    n = ctypes.c_int(0)
    rc = lib().kdejs_device_count(ctypes.byref(n))
    return n.value if rc == KDEJS_OK else 0


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device(device: int) -> int:
    """Pin the calling process to the CPUs of the NUMA node `device` hangs off, so that pinned staging
    memory allocated afterwards (first touch) and the copy-engine traffic stay on the GPU's own socket.
    Meant for one-process-per-GPU launches; returns the node, or -1 when nothing was changed (platform
    does not expose the node, KDEJS_NUMA_BIND=0, or the affinity call is not permitted)."""
    import os
    if os.environ.get("KDEJS_NUMA_BIND", "1") == "0":
        return -1
    node = ctypes.c_int(-1)
    if lib().kdejs_device_numa_node(int(device), ctypes.byref(node)) != KDEJS_OK or node.value < 0:
        return -1
    try:
        with open(f"/sys/devices/system/node/node{node.value}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return -1
        os.sched_setaffinity(0, cpus)
    except (OSError, ValueError):
        return -1
    return node.value


def as_points(x, name="array") -> np.ndarray:
    """(N,2) C-contiguous float64 copy-or-view; never mutates the caller's array."""
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError(f"Expected 2D array, got {a.ndim}D array instead ({name})")
    if a.shape[1] != 2:
        raise ValueError(f"{name} must have exactly 2 columns (core, accessory); got shape {a.shape}")
    if a.shape[0] == 0:
        raise ValueError(f"Found array with 0 sample(s) (shape={a.shape}) while a minimum of 1 is required.")
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pointer_array(arrays):
    arr = (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    cnt = (ctypes.c_int64 * len(arrays))(*[a.shape[0] for a in arrays])
    return arr, cnt


_addressof, _c_char = ctypes.addressof, ctypes.c_char
_F8, _F8_ROW_STRIDES = np.dtype(np.float64), (16, 8)


class PointSets:
    """A batch of (N_i,2) float64 point sets as the C ABI wants it: `addr` (uint64 host addresses) and `count`
    (int64 rows), numpy arrays that `pointers()` hands to ctypes without a per-set Python loop.

    `sets` is a sequence of (N_i,2) arrays, or ONE (B,N,2) array (what an ELFI batch of equal-sized simulations
    is): then nothing is touched per set -- 64 sets cost ~5 us of Python instead of ~100."""

    def __init__(self, sets, name="sims"):
        if isinstance(sets, np.ndarray) and sets.ndim == 3:
            a = sets
            if a.shape[2] != 2:
                raise ValueError(f"{name} must have exactly 2 columns (core, accessory); got shape {a.shape}")
            if a.shape[0] and a.shape[1] == 0:
                raise ValueError(f"Found array with 0 sample(s) (shape={a.shape[1:]}) while a minimum of 1 is required.")
            if a.dtype != np.float64 or a.strides[2] != 8 or a.strides[1] != 16 or a.strides[0] < 0:
                a = np.ascontiguousarray(a, dtype=np.float64)
            self.keep = a
            self.addr = np.uint64(a.ctypes.data) + np.arange(a.shape[0], dtype=np.uint64) * np.uint64(a.strides[0])
            self.count = np.full(a.shape[0], a.shape[1], dtype=np.int64)
        else:
            nd, ok = np.ndarray, _F8_ROW_STRIDES        # already a C-contiguous (N,2) float64 array: taken as it is
            arrays = [s if (type(s) is nd and s.dtype is _F8 and s.strides == ok and s.shape[1] == 2 and s.shape[0] > 0)
                      else as_points(s, f"{name}[{i}]") for i, s in enumerate(sets)]
            self.keep = arrays
            try:        # cheapest way to an ndarray's address from Python (0.36 us; .ctypes.data costs 0.9)
                addr = [_addressof(_c_char.from_buffer(a)) for a in arrays]
            except (TypeError, ValueError):       # a read-only array in the list
                addr = [a.ctypes.data for a in arrays]
            self.addr = np.array(addr, dtype=np.uint64)
            self.count = np.array([a.shape[0] for a in arrays], dtype=np.int64)

    def __len__(self):
        return self.addr.shape[0]

    def sizes(self):
        return set(np.unique(self.count).tolist())

    def pointers(self, index=None):
        """(void**, int64*, n, keepalive) for all sets or for `index` (a slice)."""
        addr = self.addr if index is None else np.ascontiguousarray(self.addr[index])
        count = self.count if index is None else np.ascontiguousarray(self.count[index])
        return addr.ctypes.data_as(_c_dpp), count.ctypes.data_as(_c_i64p), addr.shape[0], (addr, count)


class PinnedPool:
    """numpy views over page-locked host memory from kdejs_host_alloc."""

    def __init__(self):
        self._ptrs = []

    def empty(self, shape, dtype=np.float64) -> np.ndarray:
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = lib().kdejs_host_alloc(max(nbytes, 1))
        if not p:
            raise MemoryError("kdejs_host_alloc failed")
        self._ptrs.append(p)
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def empty_batch(self, count, shape, dtype=np.float64):
        """`count` arrays of `shape` carved out of ONE page-locked arena, back to back: kdejs_discrepancy_batch
        then uploads a whole wave of them with a single copy instead of one per array."""
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        arena = self.empty((int(count),) + shape, dtype)
        return [arena[i] for i in range(int(count))]

    def close(self):
        for p in self._ptrs:
            lib().kdejs_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _TsvMemory:
    """Owns one kdejs_tsv_load allocation; numpy arrays built from it keep it alive (array.base)."""

    def __init__(self, ptr, rows, pinned):
        self._ptr, self._rows, self._pinned = ptr, rows, pinned
        self.__array_interface__ = {"shape": (rows, 2), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            if self._ptr:
                lib().kdejs_tsv_free(self._ptr, int(self._pinned))
                self._ptr = None
        except Exception:
            pass


class TsvTable:
    """(rows,2) float64 table read by kdejs_tsv_load (pinned memory when a GPU is present).  `.array`
    stays valid for as long as any reference to it exists, independently of this object."""

    def __init__(self, path, skip_rows=-1, pinned=None, threads=0):
        if pinned is None:
            pinned = device_count() > 0
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        rc = lib().kdejs_tsv_load(os.fsencode(path), int(skip_rows), int(bool(pinned)), int(threads),
                                  ctypes.byref(p), ctypes.byref(n))
        if rc != KDEJS_OK:
            msg = (lib().kdejs_io_last_error() or b"").decode("utf-8", "replace")
            raise (ValueError if rc == KDEJS_ERR_ARG else RuntimeError)(msg)
        self.rows, self.pinned = n.value, bool(pinned)
        self.array = np.asarray(_TsvMemory(p.value, n.value, bool(pinned)))

    def close(self):
        self.array = None       # the memory is released when the last array referencing it goes away


def tsv_load_into(path, dst, skip_rows=-1, threads=1):
    """Parse a table into the (capacity,2) float64 array `dst` (kdejs_tsv_load_into).  Returns the number of rows;
    when they do not fit, nothing is written and the NEGATED row count comes back."""
    if dst.dtype != np.float64 or dst.ndim != 2 or dst.shape[1] != 2 or not dst.flags.c_contiguous:
        raise ValueError("dst must be a C-contiguous (capacity,2) float64 array")
    n = ctypes.c_int64()
    rc = lib().kdejs_tsv_load_into(os.fsencode(path), int(skip_rows), int(threads), dst.ctypes.data, dst.shape[0],
                                   ctypes.byref(n))
    if rc == KDEJS_ERR_CAPACITY:
        return -n.value
    if rc != KDEJS_OK:
        msg = (lib().kdejs_io_last_error() or b"").decode("utf-8", "replace")
        raise (ValueError if rc == KDEJS_ERR_ARG else RuntimeError)(msg)
    return n.value


def _arena_memory(shape, dtype, pinned):
    """(pool or None, array): page-locked memory when asked for and available, ordinary memory otherwise (the library
    stages pageable uploads itself -- slower, never wrong)."""
    if pinned:
        pool = PinnedPool()
        try:
            return pool, pool.empty(shape, dtype)
        except MemoryError:
            pool.close()
    return None, np.empty(shape, dtype=dtype)


class TableArena:
    """`slots` tables of up to `capacity_rows` rows each in ONE (slots, capacity_rows, 2) float64 array (page-locked
    when a GPU is present): files are parsed straight into their slot by a pool of threads (ctypes releases the
    GIL), nothing is allocated per file, and tables of equal length go to the batch entry points as one 3-D view."""

    def __init__(self, slots, capacity_rows, pinned=None):
        if pinned is None:
            pinned = device_count() > 0
        shape = (int(slots), int(capacity_rows), 2)
        self._pool, self.array = _arena_memory(shape, np.float64, pinned)
        self.rows = np.zeros(int(slots), dtype=np.int64)
        self.pinned = self._pool is not None

    def load(self, paths, skip_rows=-1, workers=8, executor=None, overflow="raise"):
        """Parse paths[k] into slot k.  Returns the tables as views: a (len(paths), rows, 2) array when all have the
        same number of rows, else a list of (rows_k, 2) views.  A table that does not fit its slot raises, or with
        overflow="alloc" is read into memory of its own (its entry of the returned list is then not a view)."""
        from concurrent.futures import ThreadPoolExecutor

        paths = list(paths)
        if len(paths) > self.array.shape[0]:
            raise ValueError(f"{len(paths)} tables for {self.array.shape[0]} slots")
        spilled = {}

        def one(k):
            n = tsv_load_into(paths[k], self.array[k], skip_rows, 1)
            if n < 0:
                if overflow != "alloc":
                    raise ValueError(f"{paths[k]}: {-n} rows do not fit a slot of {self.array.shape[1]} rows")
                spilled[k] = TsvTable(paths[k], skip_rows=skip_rows, pinned=self.pinned, threads=1).array
                n = -1
            self.rows[k] = n

        if executor is not None:
            list(executor.map(one, range(len(paths))))
        else:
            with ThreadPoolExecutor(max_workers=max(1, int(workers))) as ex:
                list(ex.map(one, range(len(paths))))
        r = self.rows[:len(paths)]
        if len(paths) and not spilled and (r == r[0]).all():
            return self.array[:len(paths), :int(r[0])]
        return [spilled[k] if k in spilled else self.array[k, :int(r[k])] for k in range(len(paths))]

    def close(self):
        self.array = None
        if self._pool is not None:
            self._pool.close()
            self._pool = None


class TextArena:
    """The BYTES of up to `slots` files in one (slots, slot_bytes) uint8 array (page-locked when a GPU is present), read
    by kdejs_files_read -- plain read(2), `threads` files at a time, no parsing: the device parses
    (KDE_distance.KDE_JS_divergence_batch_text)."""

    def __init__(self, slots, slot_bytes, pinned=None):
        if pinned is None:
            pinned = device_count() > 0
        self._pool, self.array = _arena_memory((int(slots), int(slot_bytes)), np.uint8, pinned)
        self.nbytes = np.zeros(int(slots), dtype=np.int64)

    def read(self, paths, threads=8):
        """Reads paths[k] into slot k.  self.nbytes[k] = bytes read, -(file size) if the file does not fit its slot,
        -1 if it cannot be opened."""
        paths = [os.fsencode(p) for p in paths]
        if len(paths) > self.array.shape[0]:
            raise ValueError(f"{len(paths)} files for {self.array.shape[0]} slots")
        arr = (ctypes.c_char_p * len(paths))(*paths)
        check(lib().kdejs_files_read(arr, len(paths), self.array.ctypes.data, self.array.shape[1],
                                     self.nbytes.ctypes.data_as(_c_i64p), int(threads)))
        return self.nbytes[:len(paths)]

    def close(self):
        self.array = None
        if self._pool is not None:
            self._pool.close()
            self._pool = None


def tsv_parse_text(text, skip_rows=-1, device=0):
    """The rows of a table held as bytes, parsed by the DEVICE parser (kdejs_tsv_parse_text): ((rows,2) float64,
    status) with status KDEJS_TABLE_OK or KDEJS_TABLE_HOST (something only the host reader covers: ignore the rows)."""
    buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    n, st = ctypes.c_int64(), ctypes.c_int32()
    out = np.empty((buf.shape[0] // 2 + 1, 2))          # a data line is at least one byte and a newline
    check(lib().kdejs_tsv_parse_text(int(device), buf.ctypes.data, buf.shape[0], int(skip_rows), out.ctypes.data,
                                     out.shape[0], ctypes.byref(n), ctypes.byref(st)))
    return out[:n.value].copy(), st.value


def load_numbers(path) -> np.ndarray:
    """Every number of a small text file as a float64 vector (kdejs_numbers_load; np.loadtxt(path).reshape(-1) for
    the simulator's gene-frequency files, without the GIL)."""
    p, n = ctypes.c_void_p(), ctypes.c_int64()
    rc = lib().kdejs_numbers_load(os.fsencode(path), ctypes.byref(p), ctypes.byref(n))
    if rc != KDEJS_OK:
        msg = (lib().kdejs_io_last_error() or b"").decode("utf-8", "replace")
        raise (ValueError if rc == KDEJS_ERR_ARG else RuntimeError)(msg)
    try:
        return np.ctypeslib.as_array(ctypes.cast(p, _c_dp), shape=(n.value,)).copy()
    finally:
        lib().kdejs_tsv_free(p, 0)
