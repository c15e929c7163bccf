"""Drop-in for the reference's Python module `instant_distance` (instant-distance-py/src/lib.rs:18-357), GPU backed.

Same classes, method names and argument meaning as the PyO3 module: `Config, Heuristic, Hnsw, HnswMap, Search, Neighbor`.
Everything numeric happens in libinstant_distance_b200.so through the C ABI (no CPU fallback).  Differences, all additive:
  * points may have any dimension (the reference fixes DIMENSIONS = 300, py:448, and zero-pads shorter inputs, py:367-374;
    here every point of an index is zero-padded to the longest point given at build time, and a longer query raises the
    same `TypeError("point array too long")`, py:369-370);
  * `Hnsw.search_many / HnswMap.search_many` expose the batched search the GPU is built for.
"""
import ctypes as C
import random

import numpy as np

from instant_distance_b200 import _abi

__all__ = ["Config", "Heuristic", "Hnsw", "HnswMap", "Search", "Neighbor"]


class Heuristic:
    """py:276-303.  Defaults: extend_candidates=False, keep_pruned=True (lib.rs:121-128)."""

    def __init__(self):
        self.extend_candidates = False
        self.keep_pruned = True


class Config:
    """py:216-256: Builder defaults (lib.rs:101-113); `seed` is drawn from entropy like `rand::random()`."""

    def __init__(self):
        p = _abi.default_params()
        self.ef_search = int(p.ef_search)
        self.ef_construction = int(p.ef_construction)
        self.ml = float(p.ml)
        self.seed = random.getrandbits(64)
        self.heuristic = Heuristic()

    def _params(self):
        kw = dict(ef_search=self.ef_search, ef_construction=self.ef_construction, ml=self.ml, seed=self.seed)
        if self.heuristic is None:
            kw["heuristic"] = 0
        else:
            kw.update(heuristic=1, extend_candidates=int(bool(self.heuristic.extend_candidates)),
                      keep_pruned=int(bool(self.heuristic.keep_pruned)))
        return kw


class Neighbor:
    """py:327-357."""

    def __init__(self, distance, pid, value=None):
        self.distance, self.pid, self.value = float(distance), int(pid), value

    def __repr__(self):
        if self.value is not None:
            return f"instant_distance.Neighbor(distance={self.distance}, pid={self.pid}, value={self.value!r})"
        return f"instant_distance.Item(distance={self.distance}, pid={self.pid})"


class Search:
    """py:159-209: search buffer and result set; iterate it after `index.search(point, search)`."""

    def __init__(self):
        self._ids = self._dist = None
        self._len = 0
        self._values = None
        self._cur = None

    def __iter__(self):
        return self

    def __next__(self):
        if self._cur is None or self._cur >= self._len:
            self._cur = None
            raise StopIteration
        i = self._cur
        self._cur += 1
        pid = int(self._ids[i])
        return Neighbor(self._dist[i], pid, None if self._values is None else self._values[pid])


def _to_matrix(points, dim=None):
    rows = [np.asarray(list(p), dtype=np.float32) for p in points]
    width = max([len(r) for r in rows], default=1) if dim is None else dim
    m = np.zeros((len(rows), max(width, 1)), dtype=np.float32)
    for i, r in enumerate(rows):
        if len(r) > m.shape[1]:
            raise TypeError("point array too long")
        m[i, :len(r)] = r
    return m


class Hnsw:
    """py:97-157."""

    def __init__(self, index, values=None):
        self._ix = index
        self._values = values
        info = index.info()
        self._dim, self._ef = int(info.dim), int(info.ef_search)

    @staticmethod
    def build(points, config):
        m = _to_matrix(points)
        ix, ids = _abi.Index.build(m, **config._params())
        return Hnsw(ix), [int(i) for i in ids]

    def search(self, point, search):
        q = _to_matrix([point], self._dim)
        k = max(self._ef, 1)
        ids, dist, lens = self._ix.search(q, ef_search=self._ef, k=k) if self._ef else (np.zeros((1, 1), np.uint32), np.zeros((1, 1), np.float32), [0])
        search._ids, search._dist, search._len, search._values, search._cur = ids[0], dist[0], int(lens[0]), self._values, 0

    def search_many(self, points, k=10, ef_search=None):
        """Batched Hnsw::search: returns (ids [nq, k], distances [nq, k], lens [nq])."""
        q = _to_matrix(points, self._dim) if not isinstance(points, np.ndarray) else points
        return self._ix.search(q, ef_search=ef_search or self._ef, k=k)

    def dump(self, fname):
        """py:131-137: bincode layout of `Hnsw` (ef_search, points, zero, layers)."""
        self._ix.save(fname)

    @staticmethod
    def load(fname, dim=300, M=32):
        """py:121-129.  The file does not store dim / M (fixed arrays in the reference: 300 / 32)."""
        try:
            ix, _ = _abi.Index.load(fname, dim, M)
        except _abi.IdbError as e:
            if e.status == _abi.ERR_IO:
                raise OSError(str(e)) from e
            raise ValueError(f"deserialization error: {e}") from e
        return Hnsw(ix)


class HnswMap(Hnsw):
    """py:30-95: values are kept on the host, permuted to PointId order as HnswMap::new does (lib.rs:144-149)."""

    @staticmethod
    def build(points, values, config):
        m = _to_matrix(points)
        vals = [str(v) if not isinstance(v, str) else v for v in values]
        ix, ids = _abi.Index.build(m, **config._params())
        by_pid = [None] * len(ids)
        for orig, pid in enumerate(ids):
            by_pid[int(pid)] = vals[orig]
        return HnswMap(ix, by_pid)

    @property
    def values(self):
        return self._values

    def dump(self, fname):
        """py:69-75: `HnswMap { hnsw, values }` — the Hnsw body, then Vec<MapValue::String> (u32 variant 0, u64 len, utf-8)."""
        import struct

        self._ix.save(fname)
        with open(fname, "ab") as f:
            f.write(struct.pack("<Q", len(self._values)))
            for v in self._values:
                b = v.encode("utf-8")
                f.write(struct.pack("<IQ", 0, len(b)) + b)

    @staticmethod
    def load(fname, dim=300, M=32):
        """py:58-67."""
        import struct

        try:
            ix, off = _abi.Index.load(fname, dim, M)
        except _abi.IdbError as e:
            if e.status == _abi.ERR_IO:
                raise OSError(str(e)) from e
            raise ValueError(f"deserialization error: {e}") from e
        values = []
        with open(fname, "rb") as f:
            f.seek(off)
            head = f.read(8)
            if len(head) != 8:
                raise ValueError("deserialization error: no values section (an Hnsw file, not an HnswMap?)")
            (count,) = struct.unpack("<Q", head)
            for _ in range(count):
                variant, ln = struct.unpack("<IQ", f.read(12))
                if variant != 0:
                    raise ValueError(f"deserialization error: unknown MapValue variant {variant}")
                values.append(f.read(ln).decode("utf-8"))
        if len(values) != int(ix.info().n):
            raise ValueError("deserialization error: values length does not match the point count")
        return HnswMap(ix, values)
