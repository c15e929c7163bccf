"""The reference accepts ANY ef (usize, lib.rs:35-47) and ANY Point (lib.rs:780-782): parity at the wide ends of both ranges.

* ef_search / ef_construction up to 1024 run the EF_T = 32 instantiations; larger values are accepted whenever the index has
  no more points than that (admission is `rank < ef` over at most n distinct ids, so ef > n behaves exactly like ef = n).
* rows of more than 1024 elements run the long-row kernels (query in shared memory, distances over groups of 32 chunks —
  the same fmaf chains in the same order, so still bit-identical to the oracle).
"""
import numpy as np
import pytest

from tests import datagen
from tests.test_gpu_build import _graph_equal
from tests.test_gpu_search_parity import _check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from instant_distance_b200 import _abi

    assert _abi.lib().idb_device_count() >= 1
    return _abi


@pytest.mark.parametrize("ef", [600, 1000, 1024])
def test_search_parity_large_ef(abi, oracle, ef):
    pts = datagen.uniform(6000, 24, 21)
    ix, _ = oracle.build(pts, seed=4, threads=8)
    _check(abi, oracle, ix.export(), ix, datagen.uniform(120, 24, 22), ef)


def test_ef_beyond_the_point_count(abi, oracle):
    pts = datagen.uniform(700, 16, 23)
    ix, _ = oracle.build(pts, seed=5)
    for ef in (700, 5000, 1_000_000):  # every reachable point comes back, exactly as the reference's unbounded `nearest` would
        _check(abi, oracle, ix.export(), ix, datagen.uniform(40, 16, 24), ef, k=700, counters=True)


def test_ef_above_1024_on_a_large_index_is_reported(abi, oracle):
    pts = datagen.uniform(3000, 8, 25)
    ix, _ = oracle.build(pts, seed=6)
    g = ix.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    with pytest.raises(abi.IdbError) as e:
        gpu.search(datagen.uniform(4, 8, 26), ef_search=2000, k=10)
    assert e.value.status == abi.ERR_UNSUPPORTED


@pytest.mark.parametrize("dim", [1025, 1536, 2052, 4096])
def test_distance_bit_exact_long_rows(abi, oracle, dim):
    rng = np.random.default_rng(dim)
    for _ in range(3):
        a = (rng.standard_normal(dim) * 3).astype(np.float32)
        b = (rng.standard_normal(dim) * 3).astype(np.float32)
        assert abi.distance(a, b).tobytes() == oracle.l2sq(a, b).tobytes()


@pytest.mark.parametrize("n,dim,M,ef", [(2500, 1536, 32, 100), (1500, 1027, 16, 64), (1200, 2048, 32, 40), (800, 4100, 24, 100)])
def test_search_parity_long_rows(abi, oracle, n, dim, M, ef):
    pts = datagen.uniform(n, dim, 31)
    ix, _ = oracle.build(pts, seed=7, M=M, threads=8)
    _check(abi, oracle, ix.export(), ix, datagen.uniform(100, dim, 32), ef)


@pytest.mark.parametrize("n,dim,kw", [(500, 1536, {}), (400, 1100, {"M": 16}), (300, 2048, {"keep_pruned": 0})])
def test_sequential_gpu_build_equals_oracle_long_rows(abi, oracle, n, dim, kw):
    _graph_equal(abi, oracle, datagen.uniform(n, dim, 40 + n), seed=n, **kw)


def test_simple_mode_long_rows(abi, oracle):
    _graph_equal(abi, oracle, datagen.uniform(300, 1300, 3), seed=9, heuristic=0)


def test_large_ef_construction(abi, oracle):
    _graph_equal(abi, oracle, datagen.uniform(1500, 12, 77), seed=13, ef_construction=700)
