"""Fixed-depth alpha-beta on the GPU: values and tie sets bit-exact against the CPU oracle
(live) and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))


@pytest.fixture(scope="module")
def dev(sq):
    sq.init()
    yield sq
    sq.shutdown()


def boards_of(pairs, sq):
    return np.array(pairs, dtype=np.uint32).view(sq.BOARD).reshape(-1)


def test_ab1_root_golden(dev, oracle):
    sc = oracle.tables()["scores_ab"]
    by_depth = {}
    for c in G["ab1_root"]:
        by_depth.setdefault(c["depth"], []).append(c)
    for depth, cases in by_depth.items():
        b = boards_of([(c["x"], c["o"]) for c in cases], dev)
        values, bv, bm, nodes = dev.alphabeta_root(b, 1, depth, sc)
        for i, c in enumerate(cases):
            assert values[i].tolist() == c["values"], (depth, i)
            assert int(bm[i]) == c["best_mask"] and int(bv[i]) == c["best_value"]
            assert nodes[i] > 0


def test_ab2_root_golden(dev, oracle):
    sc = oracle.tables()["scores_ab"]
    for c in G["ab2_root"]:
        values, bv, bm, _ = dev.alphabeta_root(boards_of([(c["x"], c["o"])], dev), 2, c["depth"], sc)
        assert values[0].tolist() == c["values"]
        assert int(bm[0]) == c["best_mask"] and int(bv[0]) == c["best_value"]


def subtree_array(dev, rows):
    t = np.zeros(len(rows), dev.SUBTREE)
    for i, (x, o, last, ply, side, carried) in enumerate(rows):
        t[i] = (x, o, last, ply, side, 0, carried)
    return t


def test_opening3_and_opening2_golden(dev, oracle):
    sc = oracle.tables()["scores_opening"]
    for depth in (6, 7):
        cases = [c for c in G["opening3"] if c["depth"] == depth]
        t = subtree_array(dev, [(1 << c["first"], 1 << c["reply"], c["reply"], 2, 1, sc[c["first"]]) for c in cases])
        v, _ = dev.alphabeta_subtrees(t, 3, depth, sc)
        assert v.tolist() == [c["value"] for c in cases]
    for depth in (6, 8):
        cases = [c for c in G["opening2"] if c["depth"] == depth]
        t = subtree_array(dev, [(1 << c["first"], 0, c["first"], 1, -1, 0) for c in cases])
        v, _ = dev.alphabeta_subtrees(t, 4, depth, sc)
        assert v.tolist() == [c["value"] for c in cases]


@pytest.mark.parametrize("split", ["0", "1", "2", "3", ""])
def test_ab1_root_live_oracle_all_split_depths(dev, oracle, split, monkeypatch):
    """the breadth-first split depth must not change any value"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(48, marks=(10, 12, 14, 16, 18, 20), seed=77)
    if split:
        os.environ["SQ_AB_SPLIT"] = split
    else:
        os.environ.pop("SQ_AB_SPLIT", None)
    try:
        values, bv, bm, _ = dev.alphabeta_root(boards, 1, 10, sc)
    finally:
        os.environ.pop("SQ_AB_SPLIT", None)
    for i in range(boards.shape[0]):
        ov, obm, obv, _ = oracle.ab1_root(int(boards[i]["x"]), int(boards[i]["o"]), 10, sc)
        assert values[i].tolist() == ov, i
        assert (int(bm[i]), int(bv[i])) == (obm, obv)


def test_shortcuts_do_not_change_values(dev, oracle):
    """win / 3-line move classes and mate-distance bounds are exact: same values with them off,
    and on boards that CAN fill up inside the horizon (where they must switch themselves off)"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(30, marks=(10, 12, 14, 16, 18, 20, 22), seed=81)
    on = dev.alphabeta_root(boards, 1, 12, sc)
    os.environ["SQ_AB_NO_SHORTCUTS"] = "1"
    try:
        off = dev.alphabeta_root(boards, 1, 12, sc)
    finally:
        os.environ.pop("SQ_AB_NO_SHORTCUTS", None)
    assert (on[0] == off[0]).all() and (on[2] == off[2]).all()
    assert on[3].sum() < off[3].sum()            # and they do save work
    for i in range(boards.shape[0]):
        x, o = int(boards[i]["x"]), int(boards[i]["o"])
        if bin(x | o).count("1") >= 14:          # oracle at depth 12 is quick from here
            assert on[0][i].tolist() == oracle.ab1_root(x, o, 12, sc)[0], i


@pytest.mark.parametrize("budget,pool", [("1", ""), ("7", ""), ("100000000", ""), ("16", "6000"), ("", "5000")])
def test_schedule_does_not_change_values(dev, oracle, budget, pool):
    """per-pass budgets, the number of passes and running out of node pool (units then finish
    without a budget) are scheduling only: values must not move"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(24, marks=(10, 12, 14), seed=79)
    for k, v in (("SQ_AB_BUDGET", budget), ("SQ_AB_POOL_NODES", pool)):
        if v:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    try:
        values, bv, bm, _ = dev.alphabeta_root(boards, 1, 10, sc)
        v3, _ = dev.alphabeta_subtrees(subtree_array(dev, [(1 << 12, 1 << 6, 6, 2, 1, 0)]), 3, 6, oracle.tables()["scores_opening"])
    finally:
        os.environ.pop("SQ_AB_BUDGET", None)
        os.environ.pop("SQ_AB_POOL_NODES", None)
    for i in range(boards.shape[0]):
        ov, obm, obv, _ = oracle.ab1_root(int(boards[i]["x"]), int(boards[i]["o"]), 10, sc)
        assert values[i].tolist() == ov and (int(bm[i]), int(bv[i])) == (obm, obv), i
    assert int(v3[0]) == oracle.ab_subtree(3, 1 << 12, 1 << 6, 6, 2, 1, 0, 6, oracle.tables()["scores_opening"])[0]


@pytest.mark.parametrize("env", [{"SQ_AB_MIN_BUDGET": "1"}, {"SQ_AB_MIN_BUDGET": "5", "SQ_AB_NO_YBW": "1"}, {"SQ_AB_MIN_BUDGET": "100000"},
                                 {"SQ_AB_POOL_NODES": "6000"}, {"SQ_AB_MIN_BUDGET": "3", "SQ_AB_NO_FAST_HORIZON": "1"}, {"SQ_AB_PASSES": "1"}, {},
                                 {"SQ_AB_PASSES": "1", "SQ_AB_YBW": "all"}, {"SQ_AB_PASSES": "1", "SQ_AB_YBW": "2"}, {"SQ_AB_PASSES": "1", "SQ_AB_NO_KILLERS": "1"}, {"SQ_AB_PASSES": "1", "SQ_AB_DEFER": "16"},
                                 {"SQ_AB_PASSES": "1", "SQ_AB_DEFER": "32", "SQ_AB_BUDGET": "40"}])
def test_queue_schedule_does_not_change_values(dev, oracle, env):
    """the queue-driven search (one persistent kernel): how eagerly units are split for hungry lanes, whether the
    brothers wait, a node pool too small to split in, the bit-parallel horizon evaluation on or off -- and the
    pass-driven search instead -- are scheduling only: no value may move"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(24, marks=(10, 12, 14), seed=83)
    so = oracle.tables()["scores_opening"]
    rows = [(1 << 12, 1 << 6, 6, 2, 1, so[12]), (1 << 0, 1 << 24, 24, 2, 1, so[0]), (1 << 7, 0, 7, 1, -1, 0)]
    env = dict(env)
    if env and "SQ_AB_PASSES" not in env:
        env["SQ_AB_QUEUE"] = "1"                 # the queue-driven search is opt-in
    for k, v in env.items():
        os.environ[k] = v
    try:
        values, bv, bm, _ = dev.alphabeta_root(boards, 1, 10, sc)
        v2, _, _, _ = dev.alphabeta_root(boards[:8], 2, 9, sc)
        v3, _ = dev.alphabeta_subtrees(subtree_array(dev, rows[:2]), 3, 7, so)
        v4, _ = dev.alphabeta_subtrees(subtree_array(dev, rows[2:]), 4, 7, so)
    finally:
        for k in env:
            os.environ.pop(k, None)
    for i in range(boards.shape[0]):
        ov, obm, obv, _ = oracle.ab1_root(int(boards[i]["x"]), int(boards[i]["o"]), 10, sc)
        assert values[i].tolist() == ov and (int(bm[i]), int(bv[i])) == (obm, obv), (env, i)
    for i in range(8):
        assert v2[i].tolist() == oracle.ab2_root(int(boards[i]["x"]), int(boards[i]["o"]), 9, sc)[0], (env, i)
    for i, (x, o, last, ply, side, carried) in enumerate(rows):
        want = oracle.ab_subtree(3 if i < 2 else 4, x, o, last, ply, side, carried, 7, so)[0]
        assert int((v3 if i < 2 else v4)[i if i < 2 else 0]) == want, (env, i)


def test_ab5_avoid_evaluator_live_oracle(dev, oracle):
    """src/alphabeta after SetAvoid(): deltaValue2 (alphabeta.go:382-454)"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(40, marks=(6, 8, 10, 12, 14, 16), seed=80)
    changed = 0
    for depth in (0, 3, 6, 8):
        values, bv, bm, _ = dev.alphabeta_root(boards, 5, depth, sc)
        for i in range(boards.shape[0]):
            x, o = int(boards[i]["x"]), int(boards[i]["o"])
            if depth == 8 and bin(x | o).count("1") < 10:
                continue
            ov, obm, obv, _ = oracle.ab5_root(x, o, depth, sc)
            assert values[i].tolist() == ov, (depth, i)
            assert (int(bm[i]), int(bv[i])) == (obm, obv)
            changed += ov != oracle.ab1_root(x, o, depth, sc)[0]
    assert changed > 0          # the extra terms do matter on some of these positions


def test_ab6_abbook_search_live_oracle(dev, oracle):
    """src/abbook's search (abbook.go:104-123,215-275): AB-1 with boardValue+delta handed down"""
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(40, marks=(6, 8, 10, 12, 14, 16), seed=81)
    changed = 0
    for depth in (0, 1, 3, 6, 8):      # abbook's SetDepth gives 6 / 8 / 10 (abbook.go:80-90)
        values, bv, bm, _ = dev.alphabeta_root(boards, 6, depth, sc)
        for i in range(boards.shape[0]):
            x, o = int(boards[i]["x"]), int(boards[i]["o"])
            if depth == 8 and bin(x | o).count("1") < 10:
                continue
            ov, obm, obv, _ = oracle.ab6_root(x, o, depth, sc)
            assert values[i].tolist() == ov, (depth, i)
            assert (int(bm[i]), int(bv[i])) == (obm, obv)
            changed += ov != oracle.ab1_root(x, o, depth, sc)[0]
    assert changed > 0


def test_subtree_carried_out_of_range_is_refused(dev):
    t = subtree_array(dev, [(1 << 12, 1 << 6, 6, 2, 1, 20000)])
    with pytest.raises(dev.SquavaError) as e:
        dev.alphabeta_subtrees(t, 3, 6, [0] * 25)
    assert e.value.status == -5


def test_ab2_root_live_oracle(dev, oracle):
    from squava_b200 import positions
    sc = oracle.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(32, marks=(10, 12, 14, 16), seed=78)
    for depth in (1, 2, 7, 11):
        values, bv, bm, _ = dev.alphabeta_root(boards, 2, depth, sc)
        for i in range(boards.shape[0]):
            ov, obm, obv, _ = oracle.ab2_root(int(boards[i]["x"]), int(boards[i]["o"]), depth, sc)
            assert values[i].tolist() == ov, (depth, i)
            assert (int(bm[i]), int(bv[i])) == (obm, obv)


@pytest.mark.parametrize("dialect", [1, 2, 3, 4, 5, 6])
def test_subtrees_live_oracle(dev, oracle, dialect):
    from squava_b200 import positions
    rng = np.random.Generator(np.random.PCG64(dialect))
    sc = oracle.tables()["scores_opening" if dialect in (3, 4) else "scores_ab"]
    rows = []
    for marks in (9, 12, 15, 18, 21, 23, 24):
        for _ in range(6):
            x, o = positions.random_position(rng, marks)
            occ = [c for c in range(25) if (x | o) >> c & 1]
            last = int(rng.choice(occ))
            side = -1 if (x >> last) & 1 else 1
            rows.append((x, o, last, int(rng.integers(1, 4)), side, int(rng.integers(-50, 50))))
    # positions that already hold a line elsewhere, and full boards
    rows.append((0x1364d93, 0x1ffffff & ~0x1364d93, 0, 3, -1, 5))
    depth = 8
    t = subtree_array(dev, rows)
    v, _ = dev.alphabeta_subtrees(t, dialect, depth, sc)
    for i, (x, o, last, ply, side, carried) in enumerate(rows):
        ov, _ = oracle.ab_subtree(dialect, x, o, last, ply, side, carried, depth, sc)
        assert int(v[i]) == ov, (dialect, i, rows[i])


@pytest.mark.parametrize("dialect", [1, 2, 3, 4, 5, 6])
def test_extreme_bias_tables_and_carried_values(dev, oracle, dialect):
    """bias entries up to +-120 and carried values in the thousands can outrank a win score: the
    exact shortcuts must switch themselves off and the values must still match"""
    from squava_b200 import positions
    rng = np.random.Generator(np.random.PCG64(100 + dialect))
    sc = [int(v) for v in rng.integers(-120, 121, size=25)]
    rows = []
    for marks in (12, 15, 18, 21):
        for _ in range(5):
            x, o = positions.random_position(rng, marks)
            occ = [c for c in range(25) if (x | o) >> c & 1]
            last = int(rng.choice(occ))
            side = -1 if (x >> last) & 1 else 1
            rows.append((x, o, last, int(rng.integers(1, 3)), side, int(rng.integers(-6000, 6000))))
    t = subtree_array(dev, rows)
    v, _ = dev.alphabeta_subtrees(t, dialect, 7, sc)
    for i, (x, o, last, ply, side, carried) in enumerate(rows):
        assert int(v[i]) == oracle.ab_subtree(dialect, x, o, last, ply, side, carried, 7, sc)[0], (dialect, i)
    if dialect in (1, 2, 5, 6):
        boards, _ = positions.midgame_batch(12, marks=(12, 14, 16), seed=90 + dialect)
        values, _, _, _ = dev.alphabeta_root(boards, dialect, 8, sc)
        f = {1: oracle.ab1_root, 2: oracle.ab2_root, 5: oracle.ab5_root, 6: oracle.ab6_root}[dialect]
        for i in range(12):
            assert values[i].tolist() == f(int(boards[i]["x"]), int(boards[i]["o"]), 8, sc)[0], i


def test_edge_cases(dev, oracle):
    sc = oracle.tables()["scores_ab"]
    v, bv, bm, nodes = dev.alphabeta_root(np.zeros(0, dev.BOARD), 1, 6, sc)
    assert v.shape == (0, 25)
    # full board: no move
    x = 0x1364d93
    v, bv, bm, _ = dev.alphabeta_root(boards_of([(x, 0x1ffffff & ~x)], dev), 1, 6, sc)
    assert (v == dev.NO_MOVE).all() and bm[0] == 0 and bv[0] == 0
    # depth 0: every root move is a leaf
    from squava_b200 import positions
    boards, _ = positions.midgame_batch(8, marks=(8, 10), seed=5)
    for dialect in (1, 2):
        v, bv, bm, _ = dev.alphabeta_root(boards, dialect, 1 if dialect == 2 else 0, sc)
        for i in range(8):
            f = oracle.ab1_root if dialect == 1 else oracle.ab2_root
            assert v[i].tolist() == f(int(boards[i]["x"]), int(boards[i]["o"]), 1 if dialect == 2 else 0, sc)[0]
    with pytest.raises(dev.SquavaError):
        dev.alphabeta_root(boards, 3, 6, sc)          # no root search in opening2/3
    with pytest.raises(dev.SquavaError):
        dev.alphabeta_root(boards, 1, 99, sc)
