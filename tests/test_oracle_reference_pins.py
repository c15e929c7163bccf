"""Pins the CPU oracle to everything the reference's own tests assert for this path.

Reference tests restated here:
  instant-distance/tests/all.rs:9-39    `map`               exact distances + values, seed independent
  instant-distance/tests/all.rs:41-46   `random_heuristic`  recall > 97/100
  instant-distance/tests/all.rs:48-53   `random_simple`     recall > 90/100
  instant-distance-py/test/test.py:15-35 self query returns own value first
"""
import numpy as np
import pytest

from tests import datagen


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 12345, 2**63 + 5])
def test_map_exact(oracle, seed):
    pts = np.array([[i, i] for i in range(5)], dtype=np.float32)
    values = ["zero", "one", "two", "three", "four"]
    ix, ids = oracle.build(pts, seed=seed, metric=1)  # metric 1 = the test's sqrt-Euclid Point (all.rs:93-97)
    # HnswMap::new value permutation (lib.rs:144-149): values[pid] = values_in[orig]
    by_pid = [None] * 5
    for orig, pid in enumerate(ids):
        by_pid[pid] = values[orig]
    got, dist, lens = ix.search(np.array([2.0, 2.0], dtype=np.float32))
    assert lens[0] == 5
    assert dist[0][0] == np.float32(0.0) and by_pid[got[0][0]] == "two"
    for i in (1, 2):
        assert dist[0][i] == np.float32(1.4142135) and by_pid[got[0][i]] in ("one", "three")
    for i in (3, 4):
        assert dist[0][i] == np.float32(2.828427) and by_pid[got[0][i]] in ("zero", "four")


def _randomized(oracle, seed, heuristic, threads):
    rng = np.random.default_rng(seed)
    pts = rng.random((1024, 2), dtype=np.float32)
    q = rng.random(2, dtype=np.float32)
    ix, ids = oracle.build(pts, seed=seed, metric=1, heuristic=heuristic, threads=threads)
    got, _, lens = ix.search(q)
    assert lens[0] >= 100
    bf, _ = oracle.bruteforce(pts, q, 100, metric=1)
    forced = set(ids[bf[0]].tolist())
    return len(forced & set(got[0][:100].tolist()))


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("threads", [1, 4])
def test_random_heuristic_recall(oracle, seed, threads):
    assert _randomized(oracle, seed, 1, threads) > 97


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("threads", [1, 4])
def test_random_simple_recall(oracle, seed, threads):
    assert _randomized(oracle, seed, 0, threads) > 90


def test_self_query_returns_self_1024x300(oracle):
    emb = np.random.default_rng(5).random((1024, 300), dtype=np.float32)
    ix, ids = oracle.build(emb, seed=9)
    got, dist, _ = ix.search(emb[123])
    assert got[0][0] == ids[123] and dist[0][0] == 0.0


def test_layer_schedule_matches_survey_table(oracle):
    # SURVEY §8 table, computed as lib.rs:238-249 does (f32 multiply, truncating cast)
    assert oracle.layer_schedule(1024) == [1024, 295, 85]
    assert oracle.layer_schedule(10_000) == [10000, 2885, 832, 240, 69]
    assert oracle.layer_schedule(1_000_000) == [1000000, 288539, 83254, 24022, 6931, 1999, 576, 166, 47]
    assert oracle.layer_schedule(10_000_000) == [10000000, 2885390, 832547, 240222, 69313, 19999, 5770, 1664, 480, 138, 39]
    assert oracle.layer_schedule(2_000_000, 24) == [2000000, 629316, 198019, 62308, 19605, 6168, 1940, 610, 191, 60]
    assert oracle.layer_schedule(10_000, 16) == [10000, 3606, 1300, 468, 168, 60, 21]
    assert oracle.layer_schedule(5) == [5]
    assert oracle.layer_schedule(0) == []


def test_shuffle_is_a_seeded_permutation(oracle):
    a = oracle.shuffle(1000, 42)
    assert sorted(a.tolist()) == list(range(1000))
    assert (a == oracle.shuffle(1000, 42)).all()
    assert (a != oracle.shuffle(1000, 43)).any()


def test_rng_core_matches_published_vectors(oracle):
    """The PointId permutation comes from rand's SmallRng (core:214, 257-260) = xoshiro256++ seeded through SplitMix64 on 64-bit
    targets; the rand crate is not under /root/reference, so the generator is pinned to the PUBLISHED known-answer vectors:
    the xoshiro256++ reference implementation's outputs from state {1,2,3,4} (also rand's own `reference` test) and the outputs of
    Xoshiro256PlusPlus::seed_from_u64(0) (rand's seeding test).  (The range reduction of random_range stays unpinned.)"""
    ref = [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205, 9973669472204895162, 14011001112246962877,
           12406186145184390807, 15849039046786891736, 10450023813501588000]
    assert oracle.rng_kat(10, state=[1, 2, 3, 4]).tolist() == ref
    seeded = [5987356902031041503, 7051070477665621255, 6633766593972829180, 211316841551650330, 9136120204379184874, 379361710973160858,
              15813423377499357806, 15596884590815070553, 5439680534584881407, 1369371744833522710]
    assert oracle.rng_kat(10, seed=0).tolist() == seeded


def test_empty_index(oracle):
    ix, ids = oracle.build(np.zeros((0, 8), dtype=np.float32))
    got, dist, lens = ix.search(np.zeros(8, dtype=np.float32), k=4)
    assert lens[0] == 0 and ids.shape == (0,) and (got == 0xFFFFFFFF).all()


def test_canonical_distance_simd_equals_scalar(oracle):
    rng = np.random.default_rng(0)
    for dim in [1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 100, 127, 128, 129, 255, 256, 300, 511, 768, 1000]:
        for _ in range(20):
            a = (rng.standard_normal(dim) * 10).astype(np.float32)
            b = (rng.standard_normal(dim) * 10).astype(np.float32)
            assert oracle.l2sq(a, b).tobytes() == oracle.l2sq(a, b, scalar=True).tobytes()
            ref = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum())
            assert abs(float(oracle.l2sq(a, b)) - ref) <= 1e-5 * max(ref, 1e-30)


def test_reference_avx2_order_within_tolerance(oracle):
    """py/src/lib.rs:390-411 (dim 300, 8-lane FMA, then the 4-lane tail) vs the canonical order: <= 1e-4 rel."""
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = rng.random(300, dtype=np.float32)
        b = rng.random(300, dtype=np.float32)
        acc = np.zeros(8, dtype=np.float32)
        for c in range(37):
            d = a[8 * c:8 * c + 8] - b[8 * c:8 * c + 8]
            acc = (d.astype(np.float64) * d.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)  # fma
        acc4 = acc[4:] + acc[:4]
        d = a[296:] - b[296:]
        acc4 = (d.astype(np.float64) * d.astype(np.float64) + acc4.astype(np.float64)).astype(np.float32)
        acc2 = acc4[:2] + acc4[2:]
        ref = np.float32(acc2[0] + acc2[1])
        got = oracle.l2sq(a, b)
        assert abs(float(got) - float(ref)) <= 1e-4 * float(ref)


def test_sequential_build_is_deterministic(oracle):
    pts = datagen.uniform(3000, 16, 3)
    g1 = oracle.build(pts, seed=11)[0].export()
    g2 = oracle.build(pts, seed=11)[0].export()
    assert (g1.zero == g2.zero).all() and all((a == b).all() for a, b in zip(g1.upper, g2.upper))


def test_graph_invariants_threaded(oracle):
    pts = datagen.uniform(5000, 24, 4)
    ix, ids = oracle.build(pts, seed=2, threads=4)
    g = ix.export()
    n, M = 5000, 32
    assert ix.layer_counts() == oracle.layer_schedule(n)
    for row_i, row in enumerate(g.zero):
        valid = row[row != 0xFFFFFFFF]
        k = len(valid)
        assert (row[:k] != 0xFFFFFFFF).all() and (row[k:] == 0xFFFFFFFF).all()  # INVALID-terminated
        assert len(set(valid.tolist())) == k and row_i not in valid and (valid < n).all()
    for l, u in enumerate(g.upper):
        assert (u[u != 0xFFFFFFFF] < u.shape[0]).all()  # upper rows only reference nodes of that layer


def test_from_graph_roundtrip_and_search_identical(oracle):
    pts = datagen.uniform(4000, 32, 8)
    ix, _ = oracle.build(pts, seed=5)
    g = ix.export()
    ix2 = oracle.from_graph(g)
    q = datagen.uniform(50, 32, 9)
    a = ix.search(q, ef_search=64, counters=True)
    b = ix2.search(q, ef_search=64, counters=True)
    for x, y in zip(a, b):
        assert (x == y).all()


def test_ties_grid_data_search_is_well_defined(oracle):
    """Duplicate vectors / exact ties: results stay sorted by (dist, pid), unique, <= ef."""
    pts = datagen.grid_ties(3000, 3, 1)
    ix, _ = oracle.build(pts, seed=1)
    q = datagen.grid_ties(100, 3, 2)
    for ef in (1, 10, 100):
        ids, dist, lens = ix.search(q, ef_search=ef)
        for i in range(100):
            L = lens[i]
            assert L <= ef and len(set(ids[i][:L].tolist())) == L
            keys = list(zip(dist[i][:L].tolist(), ids[i][:L].tolist()))
            assert keys == sorted(keys)


def test_nan_distances_follow_ordered_float_semantics(oracle):
    """`Candidate` orders by OrderedFloat<f32> distance, then pid (types.rs:228-234).  ordered-float's documented total order — NaN is
    greater than every other value and equal to itself — is not under /root/reference, so the oracle's rendering of it is stated
    here: a query with a NaN coordinate makes every distance NaN, all keys then tie on distance, and the result is what the
    traversal reaches, ordered by PointId alone; a point row with a NaN sorts after every finite distance."""
    pts = datagen.uniform(300, 8, 1)
    ix, _ = oracle.build(pts, seed=1, threads=1)
    q = datagen.uniform(1, 8, 2)
    q[0, 3] = np.nan
    ids, dist, lens = ix.search(q, ef_search=20, k=20)
    n = int(lens[0])
    assert n > 0 and np.isnan(dist[0, :n]).all()
    assert (np.diff(ids[0, :n].astype(np.int64)) > 0).all()  # ties on distance are broken by ascending pid
    # one poisoned ROW: it can only ever be the last item of a result
    pts2 = pts.copy()
    pts2[17, 0] = np.nan
    ix2, ids_map = oracle.build(pts2, seed=1, threads=1)
    q2 = datagen.uniform(50, 8, 3)
    ids2, dist2, lens2 = ix2.search(q2, ef_search=300, k=300)
    for r in range(50):
        m = int(lens2[r])
        d = dist2[r, :m]
        finite = d[~np.isnan(d)]
        assert (np.diff(finite) >= 0).all()
        if np.isnan(d).any():
            assert np.isnan(d[-1]) and np.isnan(d).sum() == 1 and ids2[r, m - 1] == ids_map[17]
