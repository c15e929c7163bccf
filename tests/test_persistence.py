"""`.idx` files: bincode-1.3 layout of `Hnsw { ef_search, points, zero, layers }` (lib.rs:193-199) as the reference's Python
binding writes it (instant-distance-py/src/lib.rs:58-75, 121-137).  No reference-written fixture exists (parity unpinned);
these tests pin OUR reader/writer to the documented layout via an independent numpy/struct statement of it."""
import os
import struct

import numpy as np
import pytest

from tests import datagen
from tests.conftest import _has_gpu

INV = 0xFFFFFFFF


def _pack(ef, pts, zero, upper):
    """Independent statement of the layout: u64 ef | u64 N, N x dim f32 | u64 N, N x 2M u32 | u64 L, L x (u64 n_l, n_l x M u32)."""
    b = struct.pack("<QQ", ef, len(pts)) + np.ascontiguousarray(pts, "<f4").tobytes()
    b += struct.pack("<Q", len(zero)) + np.ascontiguousarray(zero, "<u4").tobytes()
    b += struct.pack("<Q", len(upper))
    for u in upper:
        b += struct.pack("<Q", len(u)) + np.ascontiguousarray(u, "<u4").tobytes()
    return b


def _abi():
    from instant_distance_b200 import _abi

    return _abi


def test_load_rejects_bad_files_without_touching_the_device(tmp_path):
    abi = _abi()
    with pytest.raises(abi.IdbError) as e:
        abi.Index.load(str(tmp_path / "missing.idx"))
    assert e.value.status == abi.ERR_IO
    p = tmp_path / "short.idx"
    p.write_bytes(b"\x00" * 11)
    with pytest.raises(abi.IdbError) as e:
        abi.Index.load(str(p))
    assert e.value.status == abi.ERR_FORMAT
    pts = np.zeros((3, 4), np.float32)
    zero = np.full((3, 64), INV, np.uint32)
    good = _pack(100, pts, zero, [])
    p.write_bytes(good[:-5])
    with pytest.raises(abi.IdbError) as e:
        abi.Index.load(str(p), dim=4)
    assert e.value.status == abi.ERR_FORMAT
    p.write_bytes(good)
    with pytest.raises(abi.IdbError) as e:  # wrong dim -> the zero-layer length does not line up
        abi.Index.load(str(p), dim=5)
    assert e.value.status == abi.ERR_FORMAT
    zero[1, 0] = 77  # dangling PointId
    p.write_bytes(_pack(100, pts, zero, []))
    with pytest.raises(abi.IdbError) as e:
        abi.Index.load(str(p), dim=4)
    assert e.value.status == abi.ERR_FORMAT


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_valid_file_reaches_the_device_check(tmp_path):
    abi = _abi()
    p = tmp_path / "ok.idx"
    p.write_bytes(_pack(100, np.zeros((3, 4), np.float32), np.full((3, 64), INV, np.uint32), []))
    with pytest.raises(abi.IdbError) as e:
        abi.Index.load(str(p), dim=4)
    assert e.value.status == abi.ERR_CUDA  # parsed fine; no CPU fallback


@pytest.mark.gpu
def test_save_matches_documented_layout_and_roundtrips(tmp_path):
    abi = _abi()
    pts = datagen.uniform(3000, 300, 1)  # the reference's DIMENSIONS = 300, M = 32
    ix, ids = abi.Index.build(pts, seed=3)
    p, zero, upper = ix.export_graph()
    f = str(tmp_path / "a.idx")
    ix.save(f)
    assert open(f, "rb").read() == _pack(100, p, zero, upper)
    ix2, off = abi.Index.load(f)  # dim 300, M 32 defaults
    assert off == os.path.getsize(f)
    q = datagen.uniform(100, 300, 2)
    a, b = ix.search(q, ef_search=100, k=10), ix2.search(q, ef_search=100, k=10)
    assert (a[0] == b[0]).all() and a[1].tobytes() == b[1].tobytes()
    p2, z2, u2 = ix2.export_graph()
    assert (p2 == p).all() and (z2 == zero).all() and all((x == y).all() for x, y in zip(u2, upper))


@pytest.mark.gpu
def test_python_module_dump_load(tmp_path):
    import random

    import instant_distance

    random.seed(3)
    emb = [[random.random() for _ in range(300)] for _ in range(600)]
    vals = [f"w{i}" for i in range(600)]
    cfg = instant_distance.Config()
    m = instant_distance.HnswMap.build(emb, vals, cfg)
    f = str(tmp_path / "m.idx")
    m.dump(f)
    m2 = instant_distance.HnswMap.load(f)
    s1, s2 = instant_distance.Search(), instant_distance.Search()
    m.search(emb[5], s1)
    m2.search(emb[5], s2)
    r1, r2 = list(s1), list(s2)
    assert [(n.pid, n.value, n.distance) for n in r1] == [(n.pid, n.value, n.distance) for n in r2] and r1[0].value == "w5"
    h, _ = instant_distance.Hnsw.build(emb, cfg)
    g = str(tmp_path / "h.idx")
    h.dump(g)
    h2 = instant_distance.Hnsw.load(g)
    h2.search(emb[7], s2)
    assert next(s2).distance == 0.0
    with pytest.raises(ValueError, match="deserialization error"):
        instant_distance.HnswMap.load(g)  # an Hnsw file has no values section
    with pytest.raises(OSError):
        instant_distance.Hnsw.load(str(tmp_path / "nope.idx"))
