"""Planning artefact for the construction kernels, checked on the CPU: Search::select_heuristic
(instant-distance/src/lib.rs:636-698) visits the candidates one by one and tests each against the rows kept so far; the
"kept-major" formulation — take the first surviving candidate as kept, test ALL later survivors against it, compact, repeat —
selects the same rows in the same order while evaluating only the distances the sequential early exit needs.  The model below
computes every distance with the KEPT row as the first argument (the register-resident "query" of a future kernel) — the
reference calls candidate.distance(kept) — so the equality also shows the canonical squared-L2 is symmetric bit for bit."""
import numpy as np
import pytest

from tests import datagen


def kept_major(oracle, P, owner_vec, cand, cap, keep_pruned=True):
    d_owner = np.array([oracle.l2sq(owner_vec, P[c]) for c in cand], dtype=np.float32)
    order = sorted(range(len(cand)), key=lambda i: (float(d_owner[i]), int(cand[i])))
    surv = [(int(cand[i]), d_owner[i]) for i in order]  # ascending (distance to the owner, pid): what `nearest` holds
    cand_sorted = [c for c, _ in surv]
    kept, pruned, n_dist = [], set(), 0
    while surv and len(kept) < cap:
        r, _ = surv.pop(0)
        kept.append(r)
        nxt = []
        for c, dc in surv:
            n_dist += 1
            if oracle.l2sq(P[r], P[c]) < dc:  # strict (lib.rs:676-679); kept row first
                pruned.add(c)
            else:
                nxt.append((c, dc))
        surv = nxt
    row = list(kept)
    if keep_pruned:  # lib.rs:687-695: the discarded candidates, in candidate order, fill the row up to 2M
        row += [c for c in cand_sorted if c in pruned][: cap - len(kept)]
    return row, n_dist


@pytest.mark.parametrize("dim,M,n_cand,keep_pruned", [(128, 32, 90, True), (128, 32, 64, False), (300, 24, 200, True), (17, 8, 40, True), (2, 4, 30, True)])
def test_kept_major_selection_equals_the_sequential_heuristic(oracle, dim, M, n_cand, keep_pruned):
    rng = np.random.default_rng(dim * 1000 + n_cand)
    P = datagen.sift_shaped(3000, dim, 5) if dim >= 17 else rng.integers(0, 6, (3000, dim)).astype(np.float32)  # small dims: a grid, many ties
    # an index over these rows (no links needed) gives access to the oracle's select_heuristic with this M
    ox = oracle.from_graph(oracle.Graph(P, np.full((3000, 2 * M), 0xFFFFFFFF, np.uint32), [], M, 100))
    for trial in range(12):
        owner = int(rng.integers(0, 3000))
        # candidates: near the owner (as in a real re-prune) plus a few random ones
        d = ((P - P[owner]) ** 2).sum(1)
        near = np.argsort(d, kind="stable")[1 : n_cand - 5 + 1]
        cand = np.unique(np.concatenate([near, rng.integers(0, 3000, 5)])).astype(np.uint32)
        cand = cand[cand != owner]
        do = np.array([oracle.l2sq(P[owner], P[c]) for c in cand], dtype=np.float32)
        order = np.lexsort((cand, do))  # ascending (distance, pid)
        want_ids, _ = ox.select_heuristic(P[owner], cand[order], keep_pruned=keep_pruned)
        got, n_dist = kept_major(oracle, P, P[owner], cand, 2 * M, keep_pruned)
        assert got == [int(x) for x in want_ids], (dim, trial)
        assert n_dist <= len(cand) * min(len(cand), 2 * M)
