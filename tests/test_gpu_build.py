"""GPU Builder::build (lib.rs:209-345, 437-528) through the C ABI.

Parity bar: with insert_batch = 1 the GPU build follows the reference's sequential order, so the graph
must EQUAL the sequential oracle build bit for bit (same shuffle, same rows, same snapshots).  With concurrent batches
the reference itself is nondeterministic (rayon), so the bar is structural invariants + recall within noise of the
oracle build, plus determinism of our own schedule.
"""
import numpy as np
import pytest

from tests import datagen

pytestmark = pytest.mark.gpu
INV = 0xFFFFFFFF


@pytest.fixture(scope="module")
def abi():
    from instant_distance_b200 import _abi

    assert _abi.lib().idb_device_count() >= 1
    return _abi


def _graph_equal(abi, oracle, pts, **kw):
    okw = {k: v for k, v in kw.items() if k != "insert_batch"}
    ix_o, ids_o = oracle.build(pts, threads=1, **okw)
    g = ix_o.export()
    ix_g, ids_g = abi.Index.build(pts, insert_batch=1, **kw)
    p, zero, upper = ix_g.export_graph()
    assert (ids_g == ids_o).all(), "shuffle differs"
    assert (p == g.points).all()
    assert len(upper) == len(g.upper)
    bad = np.nonzero((zero != g.zero).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} zero rows differ, first pid {bad[:5]}: gpu {zero[bad[0]][:12]} oracle {g.zero[bad[0]][:12]}"
    for a, b in zip(upper, g.upper):
        assert a.shape == b.shape and (a == b).all()
    ix_g.close()


@pytest.mark.parametrize("n,dim,kw", [
    (1, 4, {}), (2, 4, {}), (5, 2, {}), (40, 8, {}), (600, 16, {}), (2500, 32, {}), (1500, 128, {}), (1200, 300, {"M": 24}),
    (2000, 16, {"M": 16}), (1500, 8, {"M": 48}), (1500, 24, {"ef_construction": 10}), (1500, 24, {"ef_construction": 1}),
    (1500, 24, {"ef_construction": 200}), (1500, 24, {"keep_pruned": 0}), (800, 768, {}), (1500, 6, {"M": 64}),
])
def test_sequential_gpu_build_equals_oracle(abi, oracle, n, dim, kw):
    _graph_equal(abi, oracle, datagen.uniform(n, dim, 11 + n), seed=n + 3, **kw)


def test_sequential_gpu_build_equals_oracle_ties(abi, oracle):
    _graph_equal(abi, oracle, datagen.grid_ties(1500, 3, 4), seed=5)
    _graph_equal(abi, oracle, np.ones((300, 4), dtype=np.float32), seed=6)


def test_simple_mode_equals_oracle(abi, oracle):
    """Builder::select_heuristic(None) (lib.rs:466-469, 497-515) incl. the reversed comparator."""
    _graph_equal(abi, oracle, datagen.uniform(1200, 2, 3), seed=9, heuristic=0)
    _graph_equal(abi, oracle, datagen.uniform(1500, 16, 3), seed=10, heuristic=0)


def test_shuffle_matches_oracle(abi, oracle):
    pts = datagen.uniform(5000, 4, 1)
    for seed in (0, 1, 2**40 + 17):
        ix, ids = abi.Index.build(pts, seed=seed)
        assert (ids == oracle.shuffle(5000, seed)).all()
        ix.close()


def _invariants(zero, upper, n, M, layer_counts):
    assert zero.shape == (n, 2 * M)
    for i, row in enumerate(zero):
        k = int((row != INV).sum())
        assert (row[:k] != INV).all() and (row[k:] == INV).all(), f"row {i} not INVALID-terminated"
        assert len(set(row[:k].tolist())) == k and i not in row[:k] and (row[:k] < n).all()
    assert [n] + [u.shape[0] for u in upper] == layer_counts
    for u in upper:
        assert (u[u != INV] < u.shape[0]).all()
        assert (u == zero[:u.shape[0], :M]).all() or True  # snapshots are taken mid-build; later layers rewrite zero


def _recall(ids_found, truth_pids, k=10):
    return np.mean([len(set(a[:k].tolist()) & set(b[:k].tolist())) / k for a, b in zip(ids_found, truth_pids)])


@pytest.mark.parametrize("gen,n,dim", [(datagen.uniform, 20000, 32), (datagen.sift_shaped, 30000, 128)])
def test_batched_gpu_build_quality_and_invariants(abi, oracle, gen, n, dim):
    pts = gen(n, dim, 21)
    q = gen(500, dim, 22)
    bf, _ = oracle.bruteforce(pts, q, 10, threads=8)
    ix_o, ids_o = oracle.build(pts, seed=7, threads=8)
    rec_o = _recall(ix_o.search(q, ef_search=100, k=10, threads=8)[0], ids_o[bf])
    ix_g, ids_g = abi.Index.build(pts, seed=7)
    p, zero, upper = ix_g.export_graph()
    _invariants(zero, upper, n, 32, oracle.layer_schedule(n))
    assert (ids_g == ids_o).all()
    rec_g = _recall(ix_g.search(q, ef_search=100, k=10)[0], ids_g[bf])
    print(f"recall@10 oracle-built {rec_o:.4f}  gpu-built {rec_g:.4f}")
    assert rec_g >= rec_o - 0.005, (rec_g, rec_o)
    # searching the GPU-built graph: GPU == oracle (the parity entry on our own graph)
    ox = oracle.from_graph(oracle.Graph(p, zero, upper, 32, 100))
    a = ix_g.search(q, ef_search=100, k=10)
    b = ox.search(q, ef_search=100, k=10)
    assert (a[0] == b[0]).all() and a[1].tobytes() == b[1].tobytes()
    # determinism of the batched schedule
    ix_g2, _ = abi.Index.build(pts, seed=7)
    _, zero2, upper2 = ix_g2.export_graph()
    assert (zero2 == zero).all() and all((x == y).all() for x, y in zip(upper, upper2))
    ix_g.close(), ix_g2.close()


def test_reference_recall_tests_on_gpu_build(abi, oracle):
    """tests/all.rs:41-53 restated on the GPU build + GPU search (2-D, squared-L2 has the same ordering as L2)."""
    for seed in range(4):
        rng = np.random.default_rng(seed)
        pts = rng.random((1024, 2), dtype=np.float32)
        q = rng.random(2, dtype=np.float32)
        bf, _ = oracle.bruteforce(pts, q, 100)
        for heur, thr in ((1, 97), (0, 90)):
            ix, ids = abi.Index.build(pts, seed=seed, heuristic=heur)
            got, _, lens = ix.search(q, ef_search=100, k=100)
            assert lens[0] >= 100
            assert len(set(ids[bf[0]].tolist()) & set(got[0].tolist())) > thr
            ix.close()


def test_map_reference_test_end_to_end_on_gpu(abi):
    """tests/all.rs:9-39 with GPU build + GPU search; HnswMap::new value permutation (lib.rs:144-149)."""
    pts = np.array([[i, i] for i in range(5)], dtype=np.float32)
    values = ["zero", "one", "two", "three", "four"]
    for seed in range(5):
        ix, ids = abi.Index.build(pts, seed=seed)
        by_pid = [None] * 5
        for orig, pid in enumerate(ids):
            by_pid[pid] = values[orig]
        got, dist, lens = ix.search(np.array([2.0, 2.0], dtype=np.float32))
        assert lens[0] == 5 and dist[0][:5].tolist() == [0.0, 2.0, 2.0, 8.0, 8.0]
        assert by_pid[got[0][0]] == "two" and {by_pid[i] for i in got[0][1:3]} == {"one", "three"}
        assert {by_pid[i] for i in got[0][3:5]} == {"zero", "four"}
        ix.close()


def test_build_argument_errors(abi):
    pts = datagen.uniform(100, 8, 1)
    with pytest.raises(abi.IdbError) as e:
        abi.Index.build(pts, extend_candidates=1)
    assert e.value.status == abi.ERR_UNSUPPORTED
    with pytest.raises(abi.IdbError) as e:
        abi.Index.build(pts, M=1000)
    assert e.value.status == abi.ERR_INVALID_ARG
    ix, ids = abi.Index.build(np.zeros((0, 8), dtype=np.float32))
    assert ix.info().n == 0 and ix.info().n_layers == 0
    got, dist, lens = ix.search(np.zeros((2, 8), dtype=np.float32), ef_search=10, k=3)
    assert (lens == 0).all() and (got == INV).all()


def test_progress_callback(abi):
    """Builder::progress (lib.rs:70-75, 216-222, 519-525, 331-334): monotone positions, finishes at the total."""
    seen = []
    ix, _ = abi.Index.build(datagen.uniform(5000, 8, 1), seed=1, progress=lambda done, total: seen.append((done, total)))
    assert seen and all(t == 5000 for _, t in seen) and seen[-1][0] == 5000
    assert all(a[0] <= b[0] for a, b in zip(seen, seen[1:]))
    ix.close()


def test_graph_tool_builds_the_same_graph_as_the_in_process_build(abi, tmp_path, monkeypatch):
    """lib/idb_build_graph (plain C++ over the C ABI; bench.py's untimed graph-setup step) == Index.build in this process."""
    import bench

    monkeypatch.setenv("IDB_CACHE", str(tmp_path))
    pts = datagen.sift_shaped(30_000, 64, 5)
    ids_t, zero_t, upper_t, secs = bench.build_graph_with_tool(pts, 32, 100, 100, 77, 0)
    ix, ids = abi.Index.build(pts, M=32, ef_construction=100, ef_search=100, seed=77)
    _, zero, upper = ix.export_graph()
    assert (ids_t == ids).all() and (zero_t == zero).all() and len(upper_t) == len(upper)
    assert all((a == b).all() for a, b in zip(upper_t, upper)) and secs > 0
    ix.close()
