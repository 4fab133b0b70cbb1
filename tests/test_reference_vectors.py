"""Vectors produced by the REFERENCE'S OWN CODE (squava2.py, the Python prototype in the reference
tree that src/mcts and squavam.go were transliterated from), generated in the build container by
oracle/tools/make_ref_golden.py and committed as tests/golden/ref_squava2_vectors.json.

CPU part: the C oracle (and the Python restatement) must agree with the reference on every board
(terminal flag, move list, GetResult for both players) and its exact playout distribution must sit
within 3 sigma of the reference's own rollout frequencies.  GPU part: the same through the C ABI."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "ref_squava2_vectors.json")))


def test_oracle_matches_reference_on_every_board(oracle):
    import pyoracle as P
    n_term = n_draw = 0
    for b in V["boards"]:
        x, o = b["x"], b["o"]
        term = b["moves"] == []
        assert oracle.terminal(x, o) == int(term)
        bd = P.from_masks(x, o)
        moves, pterm = P.get_moves(bd)
        assert pterm == term and moves == b["moves"]
        if term:
            n_term += 1
            for player, key in ((1, "result_x"), (-1, "result_o")):
                want = b[key]
                if want is None:            # the reference asserts on a full board without a line;
                    n_draw += 1             # the Go code returns 0.0 there (mcts.go:386-387)
                    assert oracle.get_result(x, o, player) == 0.0 and oracle.flags(x, o) == 16
                else:
                    assert oracle.get_result(x, o, player) == want == P.get_result(bd, player)
    assert n_term > 1000 and n_draw == 8


def test_exact_distribution_within_3_sigma_of_reference_rollouts(oracle):
    for r in V["rollouts"]:
        marks = bin(r["x"] | r["o"]).count("1")
        n = r["n"]
        if marks >= 8:
            p, plies = oracle.playout_exact(r["x"], r["o"], r["to_move"])
            for got, pk in ((r["x_wins"], p[0]), (r["o_wins"], p[1]), (r["draws"], p[2])):
                assert abs(got / n - pk) <= 3.0 * (pk * (1 - pk) / n) ** 0.5 + 1e-12, (r, p)
            assert abs(r["plies"] / n - plies) < 0.03
        else:                                # DP too large: the oracle's own sampling, two-sample 3 sigma
            m = 200000
            out = oracle.playouts(r["x"], r["o"], r["to_move"], m, 99)
            pk = (out[0] + r["x_wins"]) / (m + n)
            assert abs(out[0] / m - r["x_wins"] / n) <= 3.0 * (pk * (1 - pk) * (1 / m + 1 / n)) ** 0.5
            assert abs(out[3] / m - r["plies"] / n) < 0.06


@pytest.mark.gpu
def test_gpu_classify_matches_reference(sq):
    sq.init()
    try:
        boards = np.array([(b["x"], b["o"]) for b in V["boards"]], dtype=np.uint32).view(sq.BOARD).reshape(-1)
        f, wm, _ = sq.classify(boards)
        for i, b in enumerate(V["boards"]):
            assert (f[i] != 0) == (b["moves"] == [])                   # GetMoves' terminal test
            if b["moves"] == [] and b["result_x"] is not None:
                assert (wm[i] == 1) == (b["result_x"] == 1.0) and (wm[i] == -1) == (b["result_o"] == 1.0)
            elif b["moves"] == []:
                assert wm[i] == 0 and f[i] == 16
            else:
                empties = [c for c in range(25) if not ((b["x"] | b["o"]) >> c) & 1]
                assert b["moves"] == empties
    finally:
        sq.shutdown()


@pytest.mark.gpu
def test_gpu_playouts_within_3_sigma_of_reference_rollouts(sq):
    sq.init(seed=0x5155415641)
    try:
        rs = V["rollouts"]
        boards = np.array([(r["x"], r["o"]) for r in rs], dtype=np.uint32).view(sq.BOARD).reshape(-1)
        tm = np.array([r["to_move"] for r in rs], dtype=np.int8)
        m = 10_000_000
        out = sq.playouts(boards, tm, m, 2026)
        for i, r in enumerate(rs):
            n = r["n"]
            for k, key in ((0, "x_wins"), (1, "o_wins")):
                pk = (int(out[i, k]) + r[key]) / (m + n)
                tol = 3.0 * (pk * (1 - pk) * (1 / m + 1 / n)) ** 0.5          # two-sample 3 sigma
                assert abs(int(out[i, k]) / m - r[key] / n) <= tol, (i, key)
            assert int(out[i, 2]) <= 1e-5 * m and r["draws"] == 0       # draws need one of 4 full boards: ~1e-6
            assert abs(int(out[i, 3]) / m - r["plies"] / n) < 0.03
    finally:
        sq.shutdown()


# ---- alpha-beta: vectors from the reference's squava.html (mechanically translated JavaScript) ----
H = json.load(open(os.path.join(HERE, "golden", "ref_squava_html_vectors.json")))


def _html_requests(case):
    """squava.html mymove() (lines 83-106) as dialect-2 subtree calls: (ply-0 evaluation, search)"""
    x, o, d = case["x"], case["o"], case["max_depth"]
    rows = []
    for cell, (stop, delta0, val) in sorted(case["cells"].items(), key=lambda kv: int(kv[0])):
        c = int(cell)
        rows.append((c, x | (1 << c), o, stop, delta0, val, d))
    return rows


def test_oracle_matches_squava_html(oracle):
    import pyoracle as P
    sc = H["scores"]
    checked = searched = 0
    for case in H["cases"]:
        for c, xm, om, stop, delta0, val, d in _html_requests(case):
            # deltavalue(0, i, j, 0): dialect 2 at ply 0 with max_depth 0 returns exactly that number
            v0, _ = oracle.ab_subtree(2, xm, om, c, 0, -1, 0, 0, sc)
            assert v0 == delta0, (case["x"], case["o"], c)
            assert stop == (abs(delta0) > 9000)
            if not stop:
                v, _ = oracle.ab_subtree(2, xm, om, c, 1, -1, delta0, d, sc)
                assert v == val, (case["x"], case["o"], c, d)
                if checked % 7 == 0:       # the Python restatement too (slower)
                    assert P.ab2(P.from_masks(xm, om), 1, -1, -20000, 20000, c, delta0, d, sc) == val
                searched += 1
            else:
                assert val == delta0
            checked += 1
    assert checked > 500 and searched > 400


@pytest.mark.gpu
def test_gpu_matches_squava_html(sq):
    sq.init()
    try:
        sc = np.array(H["scores"], np.int8)
        by_depth = {}
        for case in H["cases"]:
            for row in _html_requests(case):
                by_depth.setdefault(row[6], []).append(row)
        for d, rows in by_depth.items():
            t = np.zeros(len(rows), sq.SUBTREE)
            for i, (c, xm, om, stop, delta0, val, _) in enumerate(rows):
                t[i] = (xm, om, c, 0, -1, 0, 0)
            v0, _ = sq.alphabeta_subtrees(t, 2, 0, sc)
            assert v0.tolist() == [r[4] for r in rows]
            live = [r for r in rows if not r[3]]
            t = np.zeros(len(live), sq.SUBTREE)
            for i, (c, xm, om, stop, delta0, val, _) in enumerate(live):
                t[i] = (xm, om, c, 1, -1, 0, delta0)
            v, _ = sq.alphabeta_subtrees(t, 2, d, sc)
            assert v.tolist() == [r[5] for r in live], d
    finally:
        sq.shutdown()


# ---- the Go sources themselves, executed through a mechanical Go->Python translation ------------
GO = json.load(open(os.path.join(HERE, "golden", "ref_go_vectors.json")))
TABLES = json.load(open(os.path.join(HERE, "golden", "ref_tables.json")))


def _row(values):
    row = [-(2 ** 31)] * 25
    for k, v in values.items():
        row[int(k)] = v
    return row


def test_oracle_matches_go_alphabeta_and_movekeeper(oracle):
    """src/alphabeta ChooseMove + src/movekeeper under -D: every root value, the chosen move, the
    maximum and even the leaf counter (which depends on the visiting order)"""
    sc = TABLES["ab"]["scores"]
    n = 0
    for c in GO["alphabeta"]:
        f = oracle.ab5_root if c["avoid"] else oracle.ab1_root
        vals, bm, bv, leaves = f(c["x"], c["o"], c["depth"], sc)
        assert vals == _row(c["values"]), c
        assert bv == c["max"] and leaves == c["leaves"]
        assert c["move"] == min(k for k in range(25) if bm >> k & 1)       # deterministic: moves[0]
        n += 1
    assert n > 80


def test_oracle_matches_go_squavathr(oracle):
    sc = TABLES["thr"]["scores"]
    for c in GO["squavathr"]:
        vals, bm, bv, leaves = oracle.ab2_root(c["x"], c["o"], c["depth"], sc)
        assert vals == _row(c["values"]), c
        assert leaves == c["leaves"]


def test_oracle_matches_go_openings(oracle):
    sc = TABLES["op3"]["scores"]
    for d, rows in GO["opening3"].items():
        assert [(r["first"], r["reply"]) for r in rows] == oracle.opening3_subtrees()
        for r in rows:
            v, lv = oracle.ab_subtree(3, 1 << r["first"], 1 << r["reply"], r["reply"], 2, 1, r["delta"], int(d), sc)
            assert (v, lv) == (r["value"], r["leaves"]) and r["delta"] == sc[r["first"]]
    for d, rows in GO["opening2"].items():
        for r in rows:
            v, lv = oracle.ab_subtree(4, 1 << r["first"], 0, r["first"], 1, -1, 0, int(d), sc)
            assert (v, lv) == (r["value"], r["leaves"])


def test_oracle_matches_go_winner_checks(oracle):
    disagree = 0
    for b in GO["winners"]:
        x, o = b["x"], b["o"]
        assert oracle.terminal(x, o) == int(b["terminal"])
        assert oracle.mcts_find_winner(x, o) == b["winner"]
        assert oracle.ab_find_winner(x, o) == b["ab_winner"]
        assert oracle.get_result(x, o, 1) == b["result_x"] and oracle.get_result(x, o, -1) == b["result_o"]
        if not b["terminal"]:
            assert b["moves"] == [c for c in range(25) if not ((x | o) >> c) & 1]
        disagree += b["winner"] != b["ab_winner"]
    assert disagree > 0                      # the two Go checkers do differ on boards where both colours own lines


@pytest.mark.gpu
def test_gpu_matches_go_vectors(sq):
    sq.init()
    try:
        sc = np.array(TABLES["ab"]["scores"], np.int8)
        for avoid, dialect in ((False, 1), (True, 5)):
            by_depth = {}
            for c in GO["alphabeta"]:
                if c["avoid"] == avoid:
                    by_depth.setdefault(c["depth"], []).append(c)
            for d, cs in by_depth.items():
                b = np.array([(c["x"], c["o"]) for c in cs], dtype=np.uint32).view(sq.BOARD).reshape(-1)
                values, bv, bm, _ = sq.alphabeta_root(b, dialect, d, sc)
                for i, c in enumerate(cs):
                    assert values[i].tolist() == _row(c["values"]), (dialect, d, i)
                    assert int(bv[i]) == c["max"] and c["move"] == min(k for k in range(25) if int(bm[i]) >> k & 1)
        sct = np.array(TABLES["thr"]["scores"], np.int8)
        for c in GO["squavathr"]:
            b = np.array([(c["x"], c["o"])], dtype=np.uint32).view(sq.BOARD).reshape(-1)
            values, _, _, _ = sq.alphabeta_root(b, 2, c["depth"], sct)
            assert values[0].tolist() == _row(c["values"]), c
        sco = np.array(TABLES["op3"]["scores"], np.int8)
        for d, rows in GO["opening3"].items():
            t = np.zeros(len(rows), sq.SUBTREE)
            for i, r in enumerate(rows):
                t[i] = (1 << r["first"], 1 << r["reply"], r["reply"], 2, 1, 0, r["delta"])
            v, _ = sq.alphabeta_subtrees(t, 3, int(d), sco)
            assert v.tolist() == [r["value"] for r in rows]
        for d, rows in GO["opening2"].items():
            t = np.zeros(len(rows), sq.SUBTREE)
            for i, r in enumerate(rows):
                t[i] = (1 << r["first"], 0, r["first"], 1, -1, 0, 0)
            v, _ = sq.alphabeta_subtrees(t, 4, int(d), sco)
            assert v.tolist() == [r["value"] for r in rows]
        b = np.array([(w["x"], w["o"]) for w in GO["winners"]], dtype=np.uint32).view(sq.BOARD).reshape(-1)
        f, wm, wa = sq.classify(b)
        for i, w in enumerate(GO["winners"]):
            assert (f[i] != 0) == w["terminal"] and wm[i] == w["winner"] and wa[i] == w["ab_winner"]
    finally:
        sq.shutdown()
