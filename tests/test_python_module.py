"""The reference's Python tests (instant-distance-py/test/test.py:4-35) against the GPU-backed drop-in module."""
import random

import pytest

pytestmark = pytest.mark.gpu


def test_hsnw():
    import instant_distance

    random.seed(1)
    points = [[random.random() for _ in range(300)] for _ in range(1024)]
    config = instant_distance.Config()
    (hnsw, ids) = instant_distance.Hnsw.build(points, config)
    assert sorted(ids) == list(range(1024))
    p = [random.random() for _ in range(300)]
    search = instant_distance.Search()
    hnsw.search(p, search)
    got = list(search)
    assert len(got) == 100 and all(a.distance <= b.distance for a, b in zip(got, got[1:]))
    assert "instant_distance.Item(" in repr(got[0])


def test_hsnw_map():
    import instant_distance

    random.seed(2)
    the_chosen_one = 123
    embeddings = [[random.random() for _ in range(300)] for _ in range(1024)]
    values = [f"word{i}" for i in range(1024)]  # /usr/share/dict/words is not in this image
    config = instant_distance.Config()
    hnsw_map = instant_distance.HnswMap.build(embeddings, values, config)
    search = instant_distance.Search()
    hnsw_map.search(embeddings[the_chosen_one], search)
    first = next(search)
    assert first.value == values[the_chosen_one] and first.distance == 0.0
    assert "instant_distance.Neighbor(" in repr(first)


def test_point_rules():
    import instant_distance

    cfg = instant_distance.Config()
    cfg.seed = 7
    hnsw, _ = instant_distance.Hnsw.build([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [2.0]], cfg)  # short point zero-padded (py:367-374)
    s = instant_distance.Search()
    hnsw.search([2.0], s)
    assert next(s).distance == 0.0
    with pytest.raises(TypeError, match="point array too long"):  # py:369-370
        hnsw.search([1.0, 2.0, 3.0, 4.0], s)
    cfg.heuristic = None  # simple mode (py:238, lib.rs:466-469)
    hnsw2, _ = instant_distance.Hnsw.build([[float(i)] for i in range(50)], cfg)
    ids, dist, lens = hnsw2.search_many([[3.2], [40.9]], k=2)
    assert lens.tolist() == [50, 50] and dist[0][0] == pytest.approx(0.04, rel=1e-4)
