"""C oracle vs the independent Python restatement, plus invariants and SURVEY Appendix B."""
import random
import sys
import os

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import pyoracle as P  # noqa: E402

DRAW_X_MASKS = [0x1364d93, 0x1552ab5, 0x15aa955, 0x19364d9]

# SURVEY Appendix B (third opinion: the surveyor's own throw-away restatement)
APPENDIX_B = [
    (0x1080038, 0x0c00181, 9995, {1, 9}, 1086218, 0.566451115023),
    (0x0504208, 0x00280a1, -9996, {10}, 117636, 0.348634328634),
    (0x0c0090a, 0x02400d4, 9993, {12, 20}, 96827, 0.411308043808),
    (0x1408051, 0x0a21500, -9994, {2, 3, 9, 13, 14, 18, 19, 20}, 112508, 0.674947459947),
    (0x05c2408, 0x1800274, 10000, {8}, 1057, 0.251219336219),
    (0x0412825, 0x098c0c0, 9995, {3}, 2518, 0.280122655123),
]
ROW_1 = "· 9995 -10000 · · · -9994 · · 9995 -9994 -9996 -9996 -9994 -10000 -9994 -9994 -9996 -9996 · -9994 -9994 · · ·"


def random_board(rng, dense=False):
    k = rng.randint(15, 25) if dense else rng.randint(0, 25)
    cells = rng.sample(range(25), k)
    nx = (k + 1) // 2 if rng.random() < 0.8 else rng.randint(0, k)
    x = sum(1 << c for c in cells[:nx])
    o = sum(1 << c for c in cells[nx:])
    return x, o


def test_board_primitives_agree(oracle):
    rng = random.Random(1)
    disagreements = 0
    for it in range(6000):
        x, o = random_board(rng, dense=it % 2 == 0)
        bd = P.from_masks(x, o)
        assert oracle.terminal(x, o) == int(P.get_moves(bd)[1])
        wm, wa = P.mcts_find_winner(bd), P.ab_find_winner(bd)
        assert oracle.mcts_find_winner(x, o) == wm
        assert oracle.ab_find_winner(x, o) == wa
        assert oracle.flags(x, o) == P.flags(bd)
        for pjm in (1, -1):
            assert oracle.get_result(x, o, pjm) == P.get_result(bd, pjm)
        f = P.flags(bd)
        both = bool(f & 5) and bool(f & 10)
        if wm != wa:
            disagreements += 1
            assert both            # scan orders can only disagree when both colours own a line
        # GetResult is consistent with FindWinner (mcts.go:92-121 vs :348-388)
        assert P.get_result(bd, 1) == (1.0 if wm == 1 else 0.0)
        assert P.get_result(bd, -1) == (1.0 if wm == -1 else 0.0)
    assert disagreements > 0


def test_four_is_checked_before_three(oracle):
    x = 0b1111           # X owns cells 0..3: a quad (and two triplets)
    assert oracle.mcts_find_winner(x, 0) == 1 and oracle.ab_find_winner(x, 0) == 1
    assert oracle.mcts_find_winner(0b111, 0) == -1      # three in a row loses
    assert oracle.get_result(x, 0, 1) == 1.0 and oracle.get_result(x, 0, -1) == 0.0
    assert oracle.get_result(0b111, 0, 1) == 0.0 and oracle.get_result(0b111, 0, -1) == 1.0


def test_draw_boards(oracle):
    for x in DRAW_X_MASKS:
        o = 0x1ffffff & ~x
        assert bin(x).count("1") == 13
        assert oracle.flags(x, o) == 16
        assert oracle.mcts_find_winner(x, o) == 0 and oracle.ab_find_winner(x, o) == 0
        assert oracle.get_result(x, o, 1) == 0.0 and oracle.get_result(x, o, -1) == 0.0
        assert oracle.terminal(x, o) == 1


def test_classify_batch_matches_scalar(oracle):
    rng = random.Random(2)
    bs = [random_board(rng) for _ in range(500)]
    arr = np.array(bs, dtype=np.uint32)
    import oracle_lib
    f, wm, wa = oracle.classify(arr.view(oracle_lib.BOARD).reshape(-1))
    for i, (x, o) in enumerate(bs):
        assert f[i] == oracle.flags(x, o)
        assert wm[i] == oracle.mcts_find_winner(x, o) and wa[i] == oracle.ab_find_winner(x, o)


def test_exact_playout_distribution_agrees(oracle):
    rng = random.Random(3)
    import squava_b200.positions as pos
    g = np.random.Generator(np.random.PCG64(5))
    for k in (13, 14, 15, 16, 18):
        x, o = pos.random_position(g, k)
        tm = 1 if k % 2 == 0 else -1
        p, plies = oracle.playout_exact(x, o, tm)
        q = P.playout_exact(P.from_masks(x, o), tm)
        assert abs(sum(p) - 1.0) < 1e-12
        for a, b in zip(p + [plies], q):
            assert abs(a - b) < 1e-10


def test_sampled_playouts_match_exact(oracle):
    x, o = 0x1080038, 0x0c00181
    p, plies = oracle.playout_exact(x, o, 1)
    n = 200000
    out = oracle.playouts(x, o, 1, n, 42)
    assert sum(out[:3]) == n
    for k in range(3):
        sigma = (p[k] * (1 - p[k]) / n) ** 0.5
        assert abs(out[k] / n - p[k]) <= 4 * sigma + 1e-12
    assert abs(out[3] / n - plies) < 0.02


def test_appendix_b_vectors(oracle):
    sc = P.SCORES_AB
    for x, o, best, ties, leaves, px in APPENDIX_B:
        vals, bm, bv, lv = oracle.ab1_root(x, o, 10, sc)
        assert bv == best and {c for c in range(25) if bm >> c & 1} == ties and lv == leaves
        p, _ = oracle.playout_exact(x, o, 1)
        assert abs(p[0] - px) < 1e-9
    vals, *_ = oracle.ab1_root(APPENDIX_B[0][0], APPENDIX_B[0][1], 10, sc)
    want = [P.NO_MOVE if t == "·" else int(t) for t in ROW_1.split()]
    assert vals == want


def positions(marks, count, seed):
    import squava_b200.positions as pos
    g = np.random.Generator(np.random.PCG64(seed))
    return [pos.random_position(g, marks) for _ in range(count)]


@pytest.mark.parametrize("marks,depth", [(16, 6), (14, 5), (18, 10), (20, 12), (12, 4)])
def test_ab1_c_vs_python(oracle, marks, depth):
    sc = P.SCORES_AB
    for x, o in positions(marks, 3, marks * 100 + depth):
        vals, bm, bv, _ = oracle.ab1_root(x, o, depth, sc)
        pv = P.ab1_root(P.from_masks(x, o), depth, sc)
        assert vals == pv
        assert (bm, bv) == P.best_mask(pv)


@pytest.mark.parametrize("marks,depth", [(16, 6), (14, 5), (18, 10), (20, 13)])
def test_ab2_c_vs_python(oracle, marks, depth):
    sc = P.SCORES_AB
    for x, o in positions(marks, 3, marks * 100 + depth + 1):
        vals, bm, bv, _ = oracle.ab2_root(x, o, depth, sc)
        pv = P.ab2_root(P.from_masks(x, o), depth, sc)
        assert vals == pv
        assert (bm, bv) == P.best_mask(pv)


@pytest.mark.parametrize("marks,depth", [(16, 6), (14, 5), (18, 9), (10, 3)])
def test_ab5_c_vs_python(oracle, marks, depth):
    sc = P.SCORES_AB
    for x, o in positions(marks, 3, marks * 100 + depth + 5):
        vals, bm, bv, _ = oracle.ab5_root(x, o, depth, sc)
        pv = P.ab1_root(P.from_masks(x, o), depth, sc, avoid=True)
        assert vals == pv
        assert (bm, bv) == P.best_mask(pv)


def test_ab1_equals_plain_minimax(oracle):
    """SURVEY Appendix C's exactness note: the root value is the minimax value of the tree."""
    sc = P.SCORES_AB
    for x, o in positions(16, 4, 77) + positions(19, 4, 78):
        bd = P.from_masks(x, o)
        vals, *_ = oracle.ab1_root(x, o, 5, sc)
        for c in range(25):
            if bd[c] != 0:
                continue
            bd[c] = 1
            stop, v = P.delta1(bd, 0, c, 0, 5, sc)
            if not stop:
                v = P.minimax_ab1(bd, 1, -1, c, v, 5, sc)
            bd[c] = 0
            assert vals[c] == v


@pytest.mark.parametrize("dialect", [1, 2, 3, 4, 5])
def test_subtrees_c_vs_python(oracle, dialect):
    rng = random.Random(dialect)
    sc = P.SCORES_OPENING if dialect in (3, 4) else P.SCORES_AB
    for x, o in positions(15, 3, 300 + dialect) + positions(21, 3, 400 + dialect):
        bd = P.from_masks(x, o)
        occupied = [c for c in range(25) if bd[c] != 0]
        last = rng.choice(occupied)
        side = -bd[last]
        ply, md, carried = 2, 6, rng.randint(-40, 40)
        v, _ = oracle.ab_subtree(dialect, x, o, last, ply, side, carried, md, sc)
        if dialect in (1, 5):
            pv = P.ab1(list(bd), ply, side, -20000, 20000, last, carried, md, sc, avoid=dialect == 5)
        elif dialect == 2:
            pv = P.ab2(list(bd), ply, side, -20000, 20000, last, carried, md, sc)
        else:
            pv = P.ab3(list(bd), ply, side, -10000, 10000, last, carried, md, sc,
                       sentinel=10000 if dialect == 4 else 20000)
        assert v == pv


def test_full_board_sentinels(oracle):
    """A node with no legal move returns the initial -+20000 (opening2: -+10000)."""
    x = DRAW_X_MASKS[0]
    o = 0x1ffffff & ~x
    last = 0 if x & 1 else 1
    sc = P.SCORES_AB
    for dialect, want in ((2, 20000), (3, 20000), (4, 10000)):
        side = 1
        v, _ = oracle.ab_subtree(dialect, x, o, last, 3, side, 0, 9, sc)
        assert v == -want
    v, _ = oracle.ab_subtree(1, x, o, last, 3, -1, 0, 9, sc)
    assert v == 20000


def test_alternating_enumeration_is_complete_and_distinct(oracle):
    """the rank -> board enumerator behind the exhaustive classification sweep"""
    import ctypes as C
    import itertools
    L = oracle.L
    L.sqo_count_alternating.restype = C.c_int64
    L.sqo_enumerate_alternating.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    assert sum(L.sqo_count_alternating(k) for k in range(11)) == 1_177_833_846
    assert L.sqo_count_alternating(25) == 5_200_300
    for k in range(5):
        n = L.sqo_count_alternating(k)
        b = np.empty(n, np.dtype([("x", "<u4"), ("o", "<u4")]))
        L.sqo_enumerate_alternating(k, 0, n, b.ctypes.data)
        got = set(zip(b["x"].tolist(), b["o"].tolist()))
        nx = (k + 1) // 2
        exp = set()
        for cells in itertools.combinations(range(25), k):
            for xs in itertools.combinations(cells, nx):
                x = sum(1 << c for c in xs)
                exp.add((x, sum(1 << c for c in cells) & ~x))
        assert len(got) == n and got == exp
    # ranges compose: two half calls equal one whole call
    n = L.sqo_count_alternating(6)
    a = np.empty(1000, np.dtype([("x", "<u4"), ("o", "<u4")])); c = a.copy()
    L.sqo_enumerate_alternating(6, n - 1000, 1000, a.ctypes.data)
    L.sqo_enumerate_alternating(6, n - 1000, 500, c.ctypes.data)
    L.sqo_enumerate_alternating(6, n - 500, 500, c[500:].ctypes.data)
    assert (a == c).all()
