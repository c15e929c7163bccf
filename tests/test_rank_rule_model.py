"""CPU model of DESIGN.md §4's exactness argument for K1: processing a whole adjacency row at once with the *rank rule* gives the
same `nearest` list as the reference's one-entry-at-a-time `Search::push` (instant-distance/src/lib.rs:704-720) followed by the
`truncate(ef)` at the end of the row (lib.rs:612).

  sequential (reference):  for x in row:  rank = #{y in nearest : y < x};  if rank < ef: nearest.insert(rank, x)      then truncate(ef)
  parallel   (K1):         S = nearest before the row;  A = {x_j : rank_S(x_j) < ef};
                           x_j admitted  iff  rank_S(x_j) + #{i < j : x_i admitted and x_i < x_j} < ef
                           nearest' = the ef smallest of S u {admitted}

Keys are (distance, pid) pairs, unique because PointIds are (types.rs:228-234); equal distances are frequent in the cases below.
(The GPU tests check the same thing end to end against the oracle; this states the rule itself.)"""
import random


def sequential(nearest, row, ef):
    near = list(nearest)
    admitted = []
    for x in row:
        rank = sum(1 for y in near if y < x)
        if rank < ef:
            near.insert(rank, x)
            admitted.append(x)
    return near[:ef], admitted


def rank_rule(nearest, row, ef):
    S = list(nearest)
    rank_s = [sum(1 for y in S if y < x) for x in row]
    admitted = []
    for j, x in enumerate(row):
        if rank_s[j] >= ef:  # can never matter: not in A
            continue
        earlier = sum(1 for a in admitted if a < x)  # admitted entries all come from A and precede j
        if rank_s[j] + earlier < ef:
            admitted.append(x)
    return sorted(S + admitted)[:ef], admitted


def test_rank_rule_equals_sequential_push():
    rng = random.Random(1234)
    for case in range(3000):
        ef = rng.choice([1, 2, 3, 5, 8, 16, 33, 100])
        n_near = rng.randint(0, ef)  # nearest never holds more than ef entries between rows
        width = rng.randint(1, 64)
        levels = rng.choice([2, 3, 10, 1000])  # few distinct distances -> many ties
        pids = rng.sample(range(10_000), n_near + width)
        keys = [(rng.randrange(levels), p) for p in pids]
        nearest = sorted(keys[:n_near])
        row = keys[n_near:]
        want, adm_seq = sequential(nearest, row, ef)
        got, adm_par = rank_rule(nearest, row, ef)
        assert got == want, (case, ef, nearest, row)
        assert adm_par == adm_seq, (case, ef)  # the same entries enter `candidates` (they are what gets expanded later)
