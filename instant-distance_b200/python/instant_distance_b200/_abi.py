"""ctypes binding of the C ABI declared in include/instant_distance_b200.h.

This is the exact stub a Python-side maintainer of the reference binding would write; nothing here
computes anything — every call goes to libinstant_distance_b200.so, which fails loudly when no CUDA device exists.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get("IDB_LIB_PATH") or os.path.join(_PKG, "lib", "libinstant_distance_b200.so")  # (IDB_LIB_PATH: A/B builds)

INVALID = 0xFFFFFFFF

OK, ERR_INVALID_ARG, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_IO, ERR_FORMAT, ERR_CAPACITY, ERR_UNSUPPORTED = range(9)

SYMBOLS = [
    "idb_params_default", "idb_build_f32", "idb_index_from_graph_f32", "idb_index_from_graph_bf16", "idb_search_batch_f32",
    "idb_search_batch_device", "idb_search_batch_device_lane", "idb_last_search_counters", "idb_last_search_failures", "idb_last_search_retried",
    "idb_index_num_lanes", "idb_index_lane_stream", "idb_device_set_persisting_l2", "idb_index_info", "idb_index_export_points",
    "idb_index_export_zero", "idb_index_export_upper", "idb_index_save", "idb_index_load", "idb_index_set_profiling", "idb_index_last_kernel_ms", "idb_debug_gather_bench", "idb_debug_gather_mix_bench",
    "idb_index_stream", "idb_index_sync", "idb_index_free",
    "idb_comm_unique_id", "idb_comm_create", "idb_comm_free", "idb_index_set_id_map", "idb_sharded_search_batch_f32",
    "idb_sharded_search_batch_device", "idb_sharded_search_batch_f32_multi", "idb_sharded_search_batch_device_multi", "idb_distance_f32", "idb_host_alloc", "idb_host_free", "idb_last_error", "idb_version", "idb_device_count",
]


class Params(C.Structure):
    _fields_ = [
        ("M", C.c_uint32), ("ef_construction", C.c_uint32), ("ef_search", C.c_uint32), ("ml", C.c_float),
        ("seed", C.c_uint64), ("heuristic", C.c_int32), ("extend_candidates", C.c_int32), ("keep_pruned", C.c_int32),
        ("insert_batch", C.c_uint32), ("device", C.c_int32), ("storage", C.c_uint32),
        ("progress", C.c_void_p), ("progress_user", C.c_void_p),
    ]


class Info(C.Structure):
    _fields_ = [
        ("n", C.c_uint64), ("dim", C.c_uint32), ("M", C.c_uint32), ("ef_search", C.c_uint32), ("n_layers", C.c_uint32),
        ("layer_n", C.c_uint64 * 32), ("device", C.c_int32), ("storage", C.c_uint32),
    ]


class IdbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"idb status {status}: {message}")
        self.status = status


_lib = None


def lib():
    """Load the shared library.  Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IdbError(ERR_CUDA, f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    u32p, f32p, u64p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.c_void_p
    L.idb_params_default.argtypes = [C.POINTER(Params)]
    L.idb_build_f32.argtypes = [f32p, C.c_uint64, C.c_uint32, C.POINTER(Params), C.POINTER(vp), u32p]
    L.idb_index_from_graph_f32.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u32p, C.c_uint32,
                                           C.POINTER(u32p), u64p, C.c_int32, C.POINTER(vp)]
    L.idb_index_from_graph_bf16.argtypes = L.idb_index_from_graph_f32.argtypes
    L.idb_search_batch_f32.argtypes = [vp, f32p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p, u32p]
    L.idb_search_batch_device.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.idb_search_batch_device_lane.argtypes = [vp, C.c_uint32, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.idb_last_search_counters.argtypes = [vp, C.c_uint64, u64p]
    L.idb_last_search_failures.argtypes = [vp, C.c_uint32, u32p]
    L.idb_last_search_retried.argtypes = [vp, C.c_uint32, u32p]
    L.idb_index_num_lanes.restype = C.c_uint32
    L.idb_index_lane_stream.argtypes = [vp, C.c_uint32]
    L.idb_index_lane_stream.restype = vp
    L.idb_device_set_persisting_l2.argtypes = [C.c_int32, C.c_int32]
    L.idb_index_info.argtypes = [vp, C.POINTER(Info)]
    L.idb_index_export_points.argtypes = [vp, f32p]
    L.idb_index_export_zero.argtypes = [vp, u32p]
    L.idb_index_export_upper.argtypes = [vp, C.c_uint32, u32p]
    L.idb_index_save.argtypes = [vp, C.c_char_p]
    L.idb_index_load.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(vp), u64p]
    L.idb_debug_gather_bench.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_double)]
    L.idb_debug_gather_mix_bench.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_double)]
    L.idb_index_set_profiling.argtypes = [vp, C.c_int32]
    L.idb_index_last_kernel_ms.argtypes = [vp, f32p, u32p]
    L.idb_index_stream.argtypes = [vp]
    L.idb_index_stream.restype = vp
    L.idb_index_sync.argtypes = [vp]
    L.idb_index_free.argtypes = [vp]
    L.idb_index_free.restype = None
    L.idb_comm_unique_id.argtypes = [vp]
    L.idb_comm_create.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.idb_comm_free.argtypes = [vp]
    L.idb_comm_free.restype = None
    L.idb_index_set_id_map.argtypes = [vp, u32p]
    L.idb_sharded_search_batch_f32.argtypes = [vp, vp, f32p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p, u32p]
    L.idb_sharded_search_batch_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.idb_sharded_search_batch_f32_multi.argtypes = [C.POINTER(vp), C.c_uint32, vp, f32p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p, u32p]
    L.idb_sharded_search_batch_device_multi.argtypes = [C.POINTER(vp), C.c_uint32, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.idb_distance_f32.argtypes = [f32p, f32p, C.c_uint32, C.c_int32, f32p]
    L.idb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.idb_host_free.argtypes = [vp]
    L.idb_host_free.restype = None
    L.idb_last_error.restype = C.c_char_p
    L.idb_version.restype = C.c_char_p
    L.idb_device_count.restype = C.c_int32
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("idb_index_stream", "idb_index_lane_stream", "idb_index_num_lanes", "idb_index_free", "idb_host_free", "idb_last_error", "idb_version", "idb_device_count",
                        "idb_comm_free"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(status):
    if status != OK:
        raise IdbError(status, lib().idb_last_error().decode("utf-8", "replace"))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


STORAGE = {"f32": 0, "bf16": 1}
PROGRESS_FN = C.CFUNCTYPE(None, C.c_uint64, C.c_uint64, C.c_void_p)


def default_params(**kw):
    p = Params()
    check(lib().idb_params_default(C.byref(p)))
    if isinstance(kw.get("storage"), str):
        kw["storage"] = STORAGE[kw["storage"]]
    if "M" in kw and "ml" not in kw:
        kw["ml"] = float(np.float32(1.0) / np.log(np.float32(kw["M"])))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Index:
    """Owns one idb_index handle."""

    def __init__(self, handle):
        self._h = handle

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.idb_index_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    @classmethod
    def from_graph(cls, points, zero, upper, M, ef_search=100, device=0, storage="f32"):
        points, zero = f32(points), np.ascontiguousarray(zero, dtype=np.uint32)
        n, dim = points.shape
        ups = [np.ascontiguousarray(u, dtype=np.uint32) for u in upper]
        arr = (C.POINTER(C.c_uint32) * max(1, len(ups)))(*[ptr(u, C.c_uint32) for u in ups])
        un = np.array([u.shape[0] for u in ups] or [0], dtype=np.uint64)
        h = C.c_void_p()
        fn = lib().idb_index_from_graph_bf16 if storage == "bf16" else lib().idb_index_from_graph_f32
        check(fn(ptr(points, C.c_float), n, dim, M, ef_search, ptr(zero, C.c_uint32), len(ups),
                                             arr, ptr(un, C.c_uint64), device, C.byref(h)))
        return cls(h)

    @classmethod
    def build(cls, rows, progress=None, **kw):
        """progress: optional callable(done, total) — Builder::progress (lib.rs:70-75)."""
        rows = f32(rows)
        n, dim = rows.shape
        p = default_params(**kw)
        cb = None
        if progress is not None:
            cb = PROGRESS_FN(lambda done, total, _user: progress(int(done), int(total)))
            p.progress = C.cast(cb, C.c_void_p)
        ids = np.empty(n, dtype=np.uint32)
        h = C.c_void_p()
        check(lib().idb_build_f32(ptr(rows, C.c_float), n, dim, C.byref(p), C.byref(h), ptr(ids, C.c_uint32)))
        return cls(h), ids

    def save(self, path):
        check(lib().idb_index_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path, dim=300, M=32, device=0):
        """Returns (Index, offset of the HnswMap values in the file)."""
        h, off = C.c_void_p(), C.c_uint64()
        check(lib().idb_index_load(os.fsencode(path), dim, M, device, C.byref(h), C.byref(off)))
        return cls(h), int(off.value)

    def info(self):
        i = Info()
        check(lib().idb_index_info(self._h, C.byref(i)))
        return i

    def _queries(self, queries):
        """n x dim f32 matrix; narrower rows are zero-padded like the reference pads short points (py:363-375), wider ones rejected."""
        q = f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        dim = int(self.info().dim)
        if q.shape[1] > dim:
            raise ValueError(f"query has {q.shape[1]} elements, the index holds {dim}-d points (py:369-370: 'point array too long')")
        if q.shape[1] < dim:
            q = np.ascontiguousarray(np.pad(q, ((0, 0), (0, dim - q.shape[1]))))
        return q

    def search(self, queries, ef_search=0, k=None):
        q = self._queries(queries)
        nq = q.shape[0]
        if k is None:
            k = ef_search or self.info().ef_search
        ids = np.empty((nq, k), dtype=np.uint32)
        dist = np.empty((nq, k), dtype=np.float32)
        lens = np.empty(nq, dtype=np.uint32)
        check(lib().idb_search_batch_f32(self._h, ptr(q, C.c_float), nq, ef_search, k, ptr(ids, C.c_uint32),
                                         ptr(dist, C.c_float), ptr(lens, C.c_uint32)))
        return ids, dist, lens

    def search_device(self, d_queries, nq, ef_search, k, d_ids, d_dist, d_len, lane=0):
        """Asynchronous: enqueues on submission lane `lane` (own stream; lanes overlap on the device)."""
        check(lib().idb_search_batch_device_lane(self._h, lane, d_queries, nq, ef_search, k, d_ids, d_dist, d_len))

    def lane_stream(self, lane):
        return lib().idb_index_lane_stream(self._h, lane)

    def last_retried(self, lane=0):
        out = C.c_uint32()
        check(lib().idb_last_search_retried(self._h, lane, C.byref(out)))
        return int(out.value)

    def last_failures(self, lane=0):
        out = C.c_uint32()
        check(lib().idb_last_search_failures(self._h, lane, C.byref(out)))
        return int(out.value)

    def set_id_map(self, global_ids):
        """global_ids[pid] = caller's id of the row that became PointId pid; None clears the map."""
        if global_ids is None:
            check(lib().idb_index_set_id_map(self._h, None))
            return
        g = np.ascontiguousarray(global_ids, dtype=np.uint32)
        if g.shape[0] != int(self.info().n):
            raise ValueError("id map must have one entry per point")
        check(lib().idb_index_set_id_map(self._h, ptr(g, C.c_uint32)))

    def sharded_search(self, comm, queries, ef_search=0, k=10):
        q = self._queries(queries)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        dist = np.empty((nq, k), dtype=np.float32)
        lens = np.empty(nq, dtype=np.uint32)
        check(lib().idb_sharded_search_batch_f32(self._h, comm._h, ptr(q, C.c_float), nq, ef_search, k, ptr(ids, C.c_uint32),
                                                 ptr(dist, C.c_float), ptr(lens, C.c_uint32)))
        return ids, dist, lens

    def sharded_search_device(self, comm, d_queries, nq, ef_search, k, d_ids, d_dist, d_len):
        check(lib().idb_sharded_search_batch_device(self._h, comm._h, d_queries, nq, ef_search, k, d_ids, d_dist, d_len))

    def last_counters(self, nq):
        out = np.zeros((nq, 4), dtype=np.uint64)
        check(lib().idb_last_search_counters(self._h, nq, ptr(out, C.c_uint64)))
        return out

    def export_graph(self):
        i = self.info()
        n, dim, M = int(i.n), int(i.dim), int(i.M)
        pts = np.empty((n, dim), dtype=np.float32)
        zero = np.empty((n, 2 * M), dtype=np.uint32)
        if n:
            check(lib().idb_index_export_points(self._h, ptr(pts, C.c_float)))
            check(lib().idb_index_export_zero(self._h, ptr(zero, C.c_uint32)))
        upper = []
        for l in range(1, int(i.n_layers)):
            u = np.empty((int(i.layer_n[l]), M), dtype=np.uint32)
            check(lib().idb_index_export_upper(self._h, l, ptr(u, C.c_uint32)))
            upper.append(u)
        return pts, zero, upper

    def gather_bench(self, n_items=10000, batches=288, chain=0, reps=3):
        ms, by = C.c_float(), C.c_double()
        check(lib().idb_debug_gather_bench(self._h, n_items, batches, chain, reps, C.byref(ms), C.byref(by)))
        return float(ms.value), float(by.value)

    def gather_mix_bench(self, n_items=10000, batches=288, chain=0, reps=3, atomics=21, mode=1):
        ms, by = C.c_float(), C.c_double()
        check(lib().idb_debug_gather_mix_bench(self._h, n_items, batches, chain, reps, atomics, mode, C.byref(ms), C.byref(by)))
        return float(ms.value), float(by.value)

    def set_profiling(self, on=True):
        check(lib().idb_index_set_profiling(self._h, 1 if on else 0))

    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_uint32()
        check(lib().idb_index_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def stream(self):
        return lib().idb_index_stream(self._h)

    def sync(self):
        check(lib().idb_index_sync(self._h))


UNIQUE_ID_BYTES = 128


def comm_unique_id():
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    check(lib().idb_comm_unique_id(buf))
    return bytes(buf.raw)


class Comm:
    """One NCCL communicator (idb_comm): rank `rank` of `world`, bound to CUDA device `device`."""

    def __init__(self, unique_id, rank, world, device):
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
        check(lib().idb_comm_create(buf, rank, world, device, C.byref(h)))
        self._h, self.rank, self.world = h, rank, world

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.idb_comm_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def _handles(shards):
    return (C.c_void_p * len(shards))(*[s._h for s in shards])


def sharded_search_multi(shards, comm, queries, ef_search=0, k=10):
    """Collective: this rank's shards (all on one device) + ONE all-gather over `comm`.  Host buffers."""
    q = shards[0]._queries(queries)
    nq = q.shape[0]
    ids = np.empty((nq, k), dtype=np.uint32)
    dist = np.empty((nq, k), dtype=np.float32)
    lens = np.empty(nq, dtype=np.uint32)
    check(lib().idb_sharded_search_batch_f32_multi(_handles(shards), len(shards), comm._h, ptr(q, C.c_float), nq, ef_search, k,
                                                   ptr(ids, C.c_uint32), ptr(dist, C.c_float), ptr(lens, C.c_uint32)))
    return ids, dist, lens


def sharded_search_multi_device(shards, comm, d_queries, nq, ef_search, k, d_ids, d_dist, d_len):
    """Device pointers; enqueues on lane 0 of shards[0] (every shard's stream is joined into it) and returns."""
    check(lib().idb_sharded_search_batch_device_multi(_handles(shards), len(shards), comm._h, d_queries, nq, ef_search, k, d_ids, d_dist, d_len))


def distance(a, b, device=0):
    a, b = f32(a), f32(b)
    out = C.c_float()
    check(lib().idb_distance_f32(ptr(a, C.c_float), ptr(b, C.c_float), a.shape[0], device, C.byref(out)))
    return np.float32(out.value)
