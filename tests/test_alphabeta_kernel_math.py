"""The identity the alpha-beta DFS kernel (squava_b200/csrc/alphabeta.cuh, ab_walk) uses for src/alphabeta's
DELAYED evaluation: a move on cell c is evaluated once per child of the node it leads to, each time with that
child's (opposing) mark j already on the board (alphabeta.go:165-223).  The kernel walks the quads through c
ONCE, keeping e1 / e2 = the empty cells that are the single hole of one / two of the mover's three-of-four quads
through c; with the extra mark on j the count of three-of-four quads is popc(e1 & ~j) + popc(e2 & ~j).
Checked against the plain count (pyoracle._lines).  No GPU needed."""
import random

import pyoracle as P


def walk(bd, c, pl):
    e1 = e2 = 0
    for q in P.IQ[c]:
        hole = [k for k in q if bd[k] != pl]
        if len(hole) == 1 and bd[hole[0]] == 0:
            bit = 1 << hole[0]
            e2 |= e1 & bit
            e1 |= bit
    return e1, e2


def test_delayed_evaluation_by_two_bit_tests():
    rng = random.Random(5)
    checked = nonzero = 0
    while checked < 60000:
        fill = rng.random()
        bd = [rng.choice((1, -1)) if rng.random() < fill else 0 for _ in range(25)]
        owned = [k for k in range(25) if bd[k] != 0]
        empty = [k for k in range(25) if bd[k] == 0]
        if not owned or not empty:
            continue
        c = rng.choice(owned)
        pl = bd[c]
        term, _ = P._lines(bd, 5, c)
        if term:
            continue                                   # the kernel keeps the walk only for non-terminal pending moves
        e1, e2 = walk(bd, c, pl)
        for j in empty:
            bd[j] = -pl
            term, v = P._lines(bd, 5, c)               # 30 per three-of-four quad through c, signed
            bd[j] = 0
            assert not term
            cnt = bin(e1 & ~(1 << j)).count("1") + bin(e2 & ~(1 << j)).count("1")
            assert v == pl * 30 * cnt, (bd, c, j)
            checked += 1
            nonzero += cnt > 0
    assert nonzero > 2000
