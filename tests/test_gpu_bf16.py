"""bf16 row storage (BASELINE config 4's data format): rows are rounded to bf16 (RNE) and kept in HBM at half the bytes;
distances still accumulate in fp32 in the canonical order.  Bar: bit-identical to the oracle run on the bf16-ROUNDED points
("oracle computes with the same bf16-rounded inputs in fp32")."""
import numpy as np
import pytest

from tests import datagen

pytestmark = pytest.mark.gpu


def bf16_round(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


@pytest.fixture(scope="module")
def abi():
    from instant_distance_b200 import _abi

    return _abi


@pytest.mark.parametrize("n,dim,M,ef", [(4000, 128, 32, 100), (3000, 300, 24, 64), (2000, 768, 32, 128), (3000, 30, 16, 200), (1200, 1536, 32, 64)])
def test_bf16_search_parity(abi, oracle, n, dim, M, ef):
    pts = datagen.uniform(n, dim, 31)
    rp = bf16_round(pts)
    ix_o, _ = oracle.build(rp, seed=4, M=M, threads=4)
    g = ix_o.export()
    q = datagen.uniform(200, dim, 32)
    for given in (g.points, ):  # rounded rows in -> stored exactly
        gpu = abi.Index.from_graph(given, g.zero, g.upper, g.M, storage="bf16")
        assert gpu.info().storage == 1
        ids, dist, lens = gpu.search(q, ef_search=ef, k=ef)
        o = ix_o.search(q, ef_search=ef, k=ef, counters=True)
        assert (ids == o[0]).all() and dist.tobytes() == o[1].tobytes() and (lens == o[2]).all()
        assert (gpu.last_counters(len(q)) == o[3]).all()
        p, _, _ = gpu.export_graph()
        assert (p == g.points).all()
        gpu.close()


def test_bf16_rounds_on_upload(abi, oracle):
    pts = datagen.uniform(1500, 64, 5)
    ix_o, ids_o = oracle.build(pts, seed=1)  # graph built on the UNROUNDED points ...
    g = ix_o.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M, storage="bf16")  # ... rows rounded by the library
    p, _, _ = gpu.export_graph()
    assert (p == bf16_round(g.points)).all() and (p != g.points).any()
    ox = oracle.from_graph(oracle.Graph(p, g.zero, g.upper, g.M, 100))
    q = datagen.uniform(100, 64, 6)
    a, b = gpu.search(q, ef_search=50, k=10), ox.search(q, ef_search=50, k=10)
    assert (a[0] == b[0]).all() and a[1].tobytes() == b[1].tobytes()


@pytest.mark.parametrize("n,dim", [(1500, 128), (1000, 768), (500, 1536)])
def test_bf16_sequential_build_equals_oracle_on_rounded_points(abi, oracle, n, dim):
    pts = datagen.uniform(n, dim, 8)
    ix_o, ids_o = oracle.build(bf16_round(pts), seed=12, threads=1)
    g = ix_o.export()
    ix_g, ids_g = abi.Index.build(pts, seed=12, insert_batch=1, storage="bf16")
    p, zero, upper = ix_g.export_graph()
    assert (ids_g == ids_o).all() and (p == g.points).all() and (zero == g.zero).all()
    assert all((a == b).all() for a, b in zip(upper, g.upper))


def test_bf16_batched_build_recall(abi, oracle):
    pts = datagen.sift_shaped(20000, 128, 3)
    q = datagen.sift_shaped(300, 128, 4)
    ix, ids = abi.Index.build(pts, seed=2, storage="bf16")
    p, _, _ = ix.export_graph()
    bf, _ = oracle.bruteforce(p, q, 10, threads=8)  # ground truth on the rounded points, PointId order
    got, _, _ = ix.search(q, ef_search=100, k=10)
    rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(got, bf)])
    assert rec > 0.97, rec
