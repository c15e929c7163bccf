"""GPU parity for the hot path: Hnsw::search (lib.rs:352-383) through the C ABI vs the CPU oracle on the same graph.

Bar (north_star): PointIds bit-identical; distances within 1e-4 relative — here they are required to be
BIT-IDENTICAL too, because both sides use the same canonical fp32 summation order; so are len(nearest) and the
traversal counters (expansions / distance evaluations per layer), which proves the traversal itself is identical.
"""
import numpy as np
import pytest

from tests import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from instant_distance_b200 import _abi

    assert _abi.lib().idb_device_count() >= 1
    return _abi


def _check(abi, oracle, g, ix_o, queries, ef, k=None, counters=True):
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M, g.ef_search)
    k = ef if k is None else k
    ids, dist, lens = gpu.search(queries, ef_search=ef, k=k)
    o_ids, o_dist, o_lens, o_cnt = ix_o.search(queries, ef_search=ef, k=k, counters=True)
    assert (lens == o_lens).all()
    assert (ids == o_ids).all(), f"{(ids != o_ids).any(axis=1).sum()} of {len(ids)} queries differ"
    assert dist.tobytes() == o_dist.tobytes()
    if counters:
        assert (gpu.last_counters(len(queries)) == o_cnt).all()
    gpu.close()


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 31, 32, 100, 127, 128, 129, 300, 768, 1000])
def test_distance_bit_exact(abi, oracle, dim):
    rng = np.random.default_rng(dim)
    for _ in range(5):
        a = (rng.standard_normal(dim) * 7).astype(np.float32)
        b = (rng.standard_normal(dim) * 7).astype(np.float32)
        assert abi.distance(a, b).tobytes() == oracle.l2sq(a, b).tobytes()


def test_map_reference_test_on_gpu(abi, oracle):
    """tests/all.rs:9-39 with the product metric (squared L2): 0, 2, 2, 8, 8 == (0, 1.4142135, 2.828427)^2."""
    pts = np.array([[i, i] for i in range(5)], dtype=np.float32)
    for seed in range(4):
        ix, ids = oracle.build(pts, seed=seed)
        g = ix.export()
        gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
        got, dist, lens = gpu.search(np.array([2.0, 2.0], dtype=np.float32))
        assert lens[0] == 5 and dist[0][:5].tolist() == [0.0, 2.0, 2.0, 8.0, 8.0]
        assert np.sqrt(dist[0][:5]).astype(np.float32).tolist() == [0.0, np.float32(1.4142135), np.float32(1.4142135),
                                                                    np.float32(2.828427), np.float32(2.828427)]
        inv = np.argsort(ids)
        assert inv[got[0][0]] == 2 and set(inv[got[0][1:3]]) == {1, 3} and set(inv[got[0][3:5]]) == {0, 4}
        assert (got[0][5:] == 0xFFFFFFFF).all() and np.isinf(dist[0][5:]).all()


@pytest.mark.parametrize("n,dim,M,ef", [
    (1, 8, 32, 10), (2, 8, 32, 10), (20, 3, 32, 100), (1024, 2, 32, 100), (3000, 32, 16, 100), (3000, 32, 24, 100),
    (5000, 128, 32, 100), (5000, 128, 32, 1), (5000, 128, 32, 7), (5000, 128, 32, 128), (4000, 300, 32, 100),
    (2000, 768, 32, 64), (3000, 16, 48, 100), (3000, 16, 64, 200), (6000, 64, 32, 300), (6000, 20, 32, 512),
])
def test_search_parity_uniform(abi, oracle, n, dim, M, ef):
    pts = datagen.uniform(n, dim, 100 + n + dim)
    ix, _ = oracle.build(pts, seed=n, M=M, threads=4)
    q = datagen.uniform(200, dim, 7)
    _check(abi, oracle, ix.export(), ix, q, ef)


def test_search_parity_10k_x32_config0(abi, oracle):
    """BASELINE.json configs[0]: 10k x 32, M=16, ef=100, 1k queries."""
    pts = datagen.uniform(10_000, 32, 42)
    ix, _ = oracle.build(pts, seed=1, M=16, threads=8)
    _check(abi, oracle, ix.export(), ix, datagen.uniform(1000, 32, 43), 100, k=10)


def test_search_parity_pid_space_beyond_16_bits(abi, oracle):
    """n > 65535: the b16 visited tables (the b16 flavour) split a PointId into (quotient, 16-bit remainder)."""
    pts = datagen.uniform(150_000, 8, 3)
    ix, _ = oracle.build(pts, seed=5, threads=8)
    _check(abi, oracle, ix.export(), ix, datagen.uniform(2000, 8, 4), 100, k=10)
    _check(abi, oracle, ix.export(), ix, datagen.uniform(500, 8, 5), 128)


def test_search_parity_sift_shaped(abi, oracle):
    pts = datagen.sift_shaped(20_000, 128, 1)
    ix, _ = oracle.build(pts, seed=3, threads=8)
    _check(abi, oracle, ix.export(), ix, datagen.sift_shaped(500, 128, 2), 100, k=10)


@pytest.mark.parametrize("ef", [1, 2, 10, 100])
@pytest.mark.parametrize("dim,side", [(3, 12), (2, 6), (8, 2)])
def test_search_parity_ties_and_duplicates(abi, oracle, ef, dim, side):
    """Integer grid: exact distance ties and duplicate vectors at the ef boundary (SURVEY §7 hard part 1)."""
    pts = datagen.grid_ties(3000, dim, 5, side=side)
    ix, _ = oracle.build(pts, seed=2)
    _check(abi, oracle, ix.export(), ix, datagen.grid_ties(300, dim, 6, side=side), ef)


def test_all_points_identical(abi, oracle):
    pts = np.ones((500, 4), dtype=np.float32)
    ix, _ = oracle.build(pts, seed=2)
    _check(abi, oracle, ix.export(), ix, np.ones((10, 4), dtype=np.float32), 10)
    _check(abi, oracle, ix.export(), ix, np.zeros((10, 4), dtype=np.float32), 100)


def test_empty_index_and_ef_zero(abi, oracle):
    gpu = abi.Index.from_graph(np.zeros((0, 8), dtype=np.float32), np.zeros((0, 64), dtype=np.uint32), [], 32)
    ids, dist, lens = gpu.search(np.zeros((3, 8), dtype=np.float32), ef_search=10, k=4)
    assert (lens == 0).all() and (ids == 0xFFFFFFFF).all() and np.isinf(dist).all()
    pts = datagen.uniform(100, 8, 1)
    ix, _ = oracle.build(pts, seed=1)
    g = ix.export(ef_search=0)
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M, 0)
    ids, dist, lens = gpu.search(pts[:3], ef_search=0, k=4)
    assert (lens == 0).all() and (ids == 0xFFFFFFFF).all()


def test_k_larger_than_result_and_smaller(abi, oracle):
    pts = datagen.uniform(50, 8, 1)
    ix, _ = oracle.build(pts, seed=1)
    g = ix.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    ids, dist, lens = gpu.search(pts[:5], ef_search=100, k=120)
    assert (lens == 50).all() and (ids[:, 50:] == 0xFFFFFFFF).all() and (ids[:, 0] != 0xFFFFFFFF).all()
    ids3, _, lens3 = gpu.search(pts[:5], ef_search=100, k=3)
    assert (ids3 == ids[:, :3]).all() and (lens3 == 50).all()


def test_self_query_returns_self_1024x300(abi, oracle):
    """instant-distance-py/test/test.py:15-35."""
    emb = np.random.default_rng(5).random((1024, 300), dtype=np.float32)
    ix, ids = oracle.build(emb, seed=9)
    g = ix.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    got, dist, _ = gpu.search(emb[123])
    assert got[0][0] == ids[123] and dist[0][0] == 0.0


def test_nan_query_orders_last(abi, oracle):
    pts = datagen.uniform(300, 8, 1)
    ix, _ = oracle.build(pts, seed=1)
    q = datagen.uniform(4, 8, 2)
    q[1, 3] = np.nan
    _check(abi, oracle, ix.export(), ix, q, 20)


def test_large_batch_and_roundtrip_export(abi, oracle):
    pts = datagen.uniform(30_000, 64, 77)
    ix, _ = oracle.build(pts, seed=4, threads=8)
    g = ix.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    p2, z2, u2 = gpu.export_graph()
    assert (p2 == g.points).all() and (z2 == g.zero).all() and all((a == b).all() for a, b in zip(u2, g.upper))
    q = datagen.uniform(20_000, 64, 78)
    ids, dist, lens = gpu.search(q, ef_search=100, k=10)
    o_ids, o_dist, o_lens = ix.search(q, ef_search=100, k=10, threads=8)
    assert (ids == o_ids).all() and dist.tobytes() == o_dist.tobytes() and (lens == o_lens).all()
    # size-independent properties: sorted by (dist, id), unique ids, idempotent
    assert (np.diff(dist, axis=1) >= 0).all()
    ids_b, dist_b, _ = gpu.search(q, ef_search=100, k=10)
    assert (ids_b == ids).all() and dist_b.tobytes() == dist.tobytes()


def test_from_graph_rejects_dangling_ids(abi):
    pts = datagen.uniform(100, 8, 1)
    zero = np.full((100, 64), 0xFFFFFFFF, dtype=np.uint32)
    zero[3, 0] = 100  # == n: out of range
    with pytest.raises(abi.IdbError) as e:
        abi.Index.from_graph(pts, zero, [], 32)
    assert e.value.status == abi.ERR_INVALID_ARG
    zero[3, 0] = 99
    up = np.full((40, 32), 0xFFFFFFFF, dtype=np.uint32)
    up[0, 0] = 40  # upper rows may only name nodes of that layer
    with pytest.raises(abi.IdbError):
        abi.Index.from_graph(pts, zero, [up], 32)
    up[0, 0] = 39
    abi.Index.from_graph(pts, zero, [up], 32).close()


def test_concurrent_callers_share_one_index(abi, oracle):
    """`Hnsw<P>: Sync` (SURVEY §8b threading): any number of threads may search one index; calls serialise inside the library."""
    import threading

    pts = datagen.uniform(20_000, 32, 9)
    ix, _ = oracle.build(pts, seed=3, threads=8)
    g = ix.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    qs = [datagen.uniform(500, 32, 100 + i) for i in range(6)]
    want = [ix.search(q, ef_search=64, k=10, threads=4) for q in qs]
    got = [None] * len(qs)

    def work(i):
        for _ in range(3):
            got[i] = gpu.search(qs[i], ef_search=64, k=10)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(qs))]
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    for w, g_ in zip(want, got):
        assert (w[0] == g_[0]).all() and w[1].tobytes() == g_[1].tobytes()


def test_visited_overflow_goes_through_the_retry_pass(abi, oracle, monkeypatch):
    """A per-warp visited table that is too small aborts the query (kQueryVisitedOverflow); the device-side retry pass
    re-runs it with a 2^18-slot table.  Results must still be the oracle's."""
    pts = datagen.uniform(20_000, 16, 13)
    ix, _ = oracle.build(pts, seed=3, threads=8)
    g = ix.export()
    q = datagen.uniform(300, 16, 14)
    want = ix.search(q, ef_search=100, k=10, counters=True)
    assert want[3][:, 3].max() > 1000  # more ids per query than 3/4 of the 1024-slot table below
    monkeypatch.setenv("IDB_VIS_SLOTS", "1024")
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    got = gpu.search(q, ef_search=100, k=10)
    assert (got[0] == want[0]).all() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
    assert (gpu.last_counters(len(q)) == want[3]).all()


FLAVOURS = {"b16": {}, "bitmap": {"IDB_VIS_TIER": "1"}, "hash": {"IDB_VIS_TIER": "0"}}


@pytest.mark.parametrize("flavour", sorted(FLAVOURS))
def test_search_parity_every_visited_flavour(abi, oracle, monkeypatch, flavour):
    """The wide-layer visited set has three exact flavours (16-bit-tag bucket set in L2 / bitmap / hash set); all must give the oracle's answer."""
    for k_, v_ in FLAVOURS[flavour].items():
        monkeypatch.setenv(k_, v_)
    for n, dim, M, ef, seed in [(5000, 128, 32, 100, 1), (3000, 16, 64, 200, 2), (4000, 300, 24, 100, 3)]:
        pts = datagen.uniform(n, dim, 900 + seed)
        ix, _ = oracle.build(pts, seed=seed, M=M, threads=4)
        _check(abi, oracle, ix.export(), ix, datagen.uniform(300, dim, 17), ef)
    pts = datagen.grid_ties(3000, 3, 5, side=12)  # exact ties and duplicate vectors
    ix, _ = oracle.build(pts, seed=2)
    _check(abi, oracle, ix.export(), ix, datagen.grid_ties(300, 3, 6, side=12), 10)


def test_bucket_set_overflow_goes_through_the_retry_pass(abi, oracle, monkeypatch):
    """A bucket set that is too small for the ids a query visits hands the query to the retry pass (hash set, 2^18 slots)."""
    pts = datagen.uniform(20_000, 16, 13)
    ix, _ = oracle.build(pts, seed=3, threads=8)
    g = ix.export()
    q = datagen.uniform(300, 16, 14)
    want = ix.search(q, ef_search=100, k=10, counters=True)
    assert want[3][:, 3].max() > 1024
    monkeypatch.setenv("IDB_B16_BYTES", "2048")  # 64 buckets x 16 slots, handed to the retry pass beyond 704 ids
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    got = gpu.search(q, ef_search=100, k=10)
    assert (got[0] == want[0]).all() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
    assert (gpu.last_counters(len(q)) == want[3]).all()


@pytest.mark.parametrize("flavour", sorted(FLAVOURS))
def test_rows_with_repeated_ids(abi, oracle, monkeypatch, flavour):
    """An adopted graph may list a PointId twice in one row (the reference's Visited then skips the second, types.rs:32-40).
    The upload-time check notices and keeps such an index off the b16 flavour (which assumes distinct ids per row)."""
    for k_, v_ in FLAVOURS[flavour].items():
        monkeypatch.setenv(k_, v_)
    pts = datagen.uniform(4000, 24, 5)
    ix, _ = oracle.build(pts, seed=7, threads=4)
    g = ix.export()
    zero = g.zero.copy()
    rng = np.random.default_rng(1)
    for r in rng.choice(len(zero), 1500, replace=False):
        cnt = int((zero[r] != 0xFFFFFFFF).sum())
        if cnt >= 3:
            i, j = rng.choice(cnt, 2, replace=False)
            zero[r, j] = zero[r, i]
    g2 = oracle.Graph(g.points, zero, g.upper, g.M, g.ef_search)
    ix2 = oracle.from_graph(g2)
    _check(abi, oracle, g2, ix2, datagen.uniform(400, 24, 6), 100)


def test_search_parity_oracle_built_200k_sift(abi, oracle):
    """BASELINE configs[1] says "search on reference-built graph": here a 200 000 x 128 sift-shaped graph built by the reference
    ALGORITHM (the oracle's threaded build, lib.rs:313-318) — not by this library — searched by both at the config's ef_search=100
    with 5 000 queries: ids, distances, len(nearest) and the per-layer traversal counters must all be bit-identical."""
    import os

    pts = datagen.sift_shaped(200_000, 128, 1)
    ix, _ = oracle.build(pts, seed=20260923, threads=min(32, os.cpu_count() or 8))
    g = ix.export()
    assert ix.num_layers == 8 and ix.layer_counts()[0] == 200_000
    _check(abi, oracle, g, ix, datagen.sift_shaped(5000, 128, 2), 100, k=100)
    _check(abi, oracle, g, ix, datagen.sift_shaped(2000, 128, 3), 200, k=10)


def test_persistent_overflows_switch_the_index_to_the_atomic_flavour(abi, oracle, monkeypatch):
    """Data whose traversals visit more ids than the b16 tables hold: every query is handed to the retry pass at first; the library
    notices (sampled read-back of the overflow tally) and serves later calls from the DRAM-resident flavour.  Results never change."""
    pts = datagen.uniform(20_000, 16, 13)
    ix, _ = oracle.build(pts, seed=3, threads=8)
    g = ix.export()
    q = datagen.uniform(400, 16, 14)
    want = ix.search(q, ef_search=100, k=10, counters=True)
    monkeypatch.setenv("IDB_B16_BYTES", "4096")  # 120 buckets: handed over beyond 1320 ids, every query here visits > 2200
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    retried = []
    for _ in range(6):
        got = gpu.search(q, ef_search=100, k=10)
        assert (got[0] == want[0]).all() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
        assert (gpu.last_counters(len(q)) == want[3]).all()
        retried.append(gpu.last_retried(0xFFFFFFFF))  # the lane the host call just used
    assert retried[0] > 40 and retried[-1] == 0, retried
