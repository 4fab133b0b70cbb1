"""The two identities negascout_fast_kernel (squava_b200/csrc/negascout.cuh) rests on, restated in Python and
checked against the oracle's plain restatement of the reference (pyoracle.ns_static / the loop of negaScout):

1. INCREMENTAL CHILD EVALUATION.  On a board without a complete line, staticValue of board + one mark
   follows from the parent's raw sum: only quads / triplets THROUGH the new cell can complete; a quad through
   it becomes three-of-four for the mover (it held two of the mover's marks and no opposing one) or stops being
   three-of-four for the opponent (it held three opposing marks); the deadly diagonals through it are recomputed.
2. HORIZON SCAN.  For a node whose children all stop, the loop of negascout.go:240-262 (running maximum, cut-off,
   re-searches that return the same static value but count a leaf) equals prefix-maximum formulas.
No GPU needed."""
import random

import numpy as np

import pyoracle as P

WIN = 10000
MULT = {}          # quad (as tuple) -> how many checkable cells it contains = how often the scan meets it
for c in P.IMPORTANT:
    for q in P.IQ[c]:
        MULT[tuple(q)] = MULT.get(tuple(q), 0) + 1


def raw_sum(bd):
    """staticValue's value before the bias fallback, on a board without complete lines"""
    v = 0
    for q, m in MULT.items():
        s = sum(bd[k] for k in q)
        if s in (3, -3):
            v += s * 10 * m
    for d in P.NS_DEADLY:
        outer, inner = bd[d[0]] + bd[d[3]], bd[d[1]] + bd[d[2]]
        if inner in (2, -2) and outer == 0:
            v -= inner // 2 * 5
    return v


def deadly_term(d, bd):
    outer, inner = bd[d[0]] + bd[d[3]], bd[d[1]] + bd[d[2]]
    return -(inner // 2) * 5 if inner in (2, -2) and outer == 0 else 0


def child_by_delta(bd, c, pl, vsum, bias, cply, scores):
    """what ns_child computes: (terminal, value from X's point of view)"""
    four = three = False
    gain = 0
    for q in P.IQ[c]:
        own = sum(1 for k in q if bd[k] == pl)
        opp = sum(1 for k in q if bd[k] == -pl)
        four |= own == 3 and opp == 0
        if (own == 2 and opp == 0) or (own == 0 and opp == 3):
            gain += 30 * MULT[tuple(q)]
    for t in P.IT[c]:
        three |= sum(1 for k in t if bd[k] == pl) == 2
    after = list(bd)
    after[c] = pl
    dd = sum(deadly_term(d, after) - deadly_term(d, bd) for d in P.NS_DEADLY if c in d)
    s = vsum + pl * gain + dd
    if four:
        return True, pl * (WIN - cply)
    if three:
        return True, -pl * (WIN - cply)
    return False, (bias + pl * scores[c]) if s == 0 else s


def test_incremental_child_evaluation_equals_static_value():
    rng = random.Random(7)
    scores = P.SCORES_AB
    checked = terminal_children = 0
    while checked < 40000:
        fill = rng.random()
        bd = [rng.choice((1, -1)) if rng.random() < fill else 0 for _ in range(25)]
        stop, _ = P.ns_static(bd, 3, 99, scores)
        if stop or 0 not in bd:
            continue                                   # the kernel only expands nodes without a complete line
        vsum = raw_sum(bd)
        bias = sum(bd[k] * scores[k] for k in range(25))
        for c in [k for k in range(25) if bd[k] == 0]:
            for pl in (1, -1):
                term, val = child_by_delta(bd, c, pl, vsum, bias, 4, scores)
                bd[c] = pl
                stop, want = P.ns_static(bd, 4, 99, scores)
                bd[c] = 0
                assert (term, val) == (stop, want), (bd, c, pl)
                checked += 1
                terminal_children += term
    assert terminal_children > 1000


def loop_at_the_horizon(cur, alpha, beta):
    """negascout.go:236-262 for a node at ply == maxDepth: every child stops, a re-search returns the same value.
    -> (score, leaves counted)"""
    score, n, leaves = -30000, beta, 0
    for v in cur:
        leaves += 1
        if v > score:
            if n == beta:
                score = v
            else:
                leaves += 1                            # the re-search: staticValue runs again, same value
                score = v
        alpha = max(alpha, score)
        if alpha >= beta:
            break
        n = alpha + 1
    return score, leaves


def scan_at_the_horizon(cur, alpha, beta):
    """the kernel's lane-parallel form"""
    cur = np.asarray(cur, dtype=np.int64)
    if cur.size == 0:
        return -30000, 0
    incl = np.maximum.accumulate(cur)
    excl = np.concatenate(([-(1 << 30)], incl[:-1]))
    before = np.maximum(alpha, excl)
    first = np.arange(cur.size) == 0
    visited = first | (before < beta)
    again = visited & ~first & (cur > excl) & (before + 1 != beta)
    last = np.nonzero(visited)[0][-1]
    return int(incl[last]), int(visited.sum() + again.sum())


def test_horizon_scan_equals_the_loop():
    rng = random.Random(9)
    for _ in range(200000):
        k = rng.randrange(0, 25)
        pool = [rng.randrange(-60, 61) for _ in range(4)] + [9990, -9990, 0]
        cur = [rng.choice(pool) for _ in range(k)]
        kind = rng.random()
        if kind < 0.5:                                 # null window, the common case
            alpha = rng.randrange(-70, 70); beta = alpha + 1
        elif kind < 0.9:
            alpha = rng.randrange(-70, 70); beta = alpha + rng.randrange(1, 60)
        else:                                          # the inverted window a re-search can be handed (beta <= alpha)
            beta = rng.randrange(-70, 70); alpha = beta + rng.randrange(0, 30)
        assert scan_at_the_horizon(cur, alpha, beta) == loop_at_the_horizon(cur, alpha, beta), (cur, alpha, beta)
