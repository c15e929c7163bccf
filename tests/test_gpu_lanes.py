"""Submission lanes and the per-device table pool: concurrent launches from several lanes / several indexes share ONE pool of
per-warp scratch tables (claimed per thread block, returned clean) and must not disturb each other's results.
`Hnsw<P>: Sync` (lib.rs:352-356): any number of threads may search one index at once."""
import threading

import numpy as np
import pytest

from tests import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from instant_distance_b200 import _abi

    assert _abi.lib().idb_device_count() >= 1
    return _abi


def test_lanes_overlap_and_agree(abi, oracle):
    import torch

    pts = datagen.sift_shaped(30_000, 64, 3)
    ix_o, _ = oracle.build(pts, seed=5, threads=8)
    g = ix_o.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    nl = abi.lib().idb_index_num_lanes()
    assert nl >= 2
    nq, k = 4000, 10
    qs = [datagen.sift_shaped(nq, 64, 100 + i) for i in range(nl)]
    want = [ix_o.search(q, ef_search=100, k=k) for q in qs]
    dq = [torch.from_numpy(q).cuda() for q in qs]
    d_ids = [torch.empty((nq, k), dtype=torch.int32, device="cuda") for _ in range(nl)]
    d_dist = [torch.empty((nq, k), dtype=torch.float32, device="cuda") for _ in range(nl)]
    d_len = [torch.empty((nq,), dtype=torch.int32, device="cuda") for _ in range(nl)]
    torch.cuda.synchronize()
    for rep in range(3):  # all lanes in flight together, repeatedly
        for l in range(nl):
            gpu.search_device(dq[l].data_ptr(), nq, 100, k, d_ids[l].data_ptr(), d_dist[l].data_ptr(), d_len[l].data_ptr(), lane=l)
    gpu.sync()
    for l in range(nl):
        assert gpu.last_failures(l) == 0
        assert (d_ids[l].cpu().numpy().view(np.uint32) == want[l][0]).all()
        assert d_dist[l].cpu().numpy().tobytes() == want[l][1].tobytes()
        assert (d_len[l].cpu().numpy().view(np.uint32) == want[l][2]).all()
    gpu.close()


def test_many_indexes_share_the_device_pool(abi, oracle):
    """Eight indexes searched from eight threads at once (the sharded layout on one GPU): one table pool, no cross-talk."""
    idx, want, qs = [], [], []
    for s in range(8):
        pts = datagen.uniform(6000, 16 + 8 * (s % 3), 50 + s)
        ix_o, _ = oracle.build(pts, seed=s, threads=4)
        g = ix_o.export()
        q = datagen.uniform(1500, pts.shape[1], 70 + s)
        idx.append(abi.Index.from_graph(g.points, g.zero, g.upper, g.M))
        want.append(ix_o.search(q, ef_search=80, k=10))
        qs.append(q)
    got = [None] * 8

    def work(i):
        for _ in range(4):
            got[i] = idx[i].search(qs[i], ef_search=80, k=10)

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for w, g_ in zip(want, got):
        assert (w[0] == g_[0]).all() and w[1].tobytes() == g_[1].tobytes() and (w[2] == g_[2]).all()
    [i.close() for i in idx]


def test_build_and_search_interleave_on_one_device(abi, oracle):
    """A build (KA claims tables from the same pool) while another index is being searched from a second thread."""
    pts = datagen.uniform(8000, 32, 9)
    ix_o, _ = oracle.build(pts, seed=2, threads=8)
    g = ix_o.export()
    gpu = abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    q = datagen.uniform(2000, 32, 10)
    want = ix_o.search(q, ef_search=100, k=10)
    stop, bad = threading.Event(), []

    def searcher():
        while not stop.is_set():
            r = gpu.search(q, ef_search=100, k=10)
            if not ((r[0] == want[0]).all() and r[1].tobytes() == want[1].tobytes()):
                bad.append(1)

    t = threading.Thread(target=searcher)
    t.start()
    built, ids = abi.Index.build(datagen.uniform(20_000, 32, 11), seed=3)
    ref, ids2 = abi.Index.build(datagen.uniform(20_000, 32, 11), seed=3)
    stop.set()
    t.join()
    assert not bad
    a, b = built.export_graph(), ref.export_graph()
    assert (ids == ids2).all() and (a[1] == b[1]).all()  # the GPU build is deterministic, whatever else runs on the device
    for x in (gpu, built, ref):
        x.close()
