"""The C oracle against the committed golden vectors (regression pin; includes SURVEY
Appendix B's six positions).  CPU only."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))


def test_ab1_root(oracle):
    sc = oracle.tables()["scores_ab"]
    for c in G["ab1_root"]:
        vals, bm, bv, lv = oracle.ab1_root(c["x"], c["o"], c["depth"], sc)
        assert (vals, bm, bv, lv) == (c["values"], c["best_mask"], c["best_value"], c["leaves"])


def test_ab2_root(oracle):
    sc = oracle.tables()["scores_ab"]
    for c in G["ab2_root"]:
        vals, bm, bv, lv = oracle.ab2_root(c["x"], c["o"], c["depth"], sc)
        assert (vals, bm, bv, lv) == (c["values"], c["best_mask"], c["best_value"], c["leaves"])


def test_openings(oracle):
    sc = oracle.tables()["scores_opening"]
    for c in G["opening3"][:40]:
        v, lv = oracle.ab_subtree(3, 1 << c["first"], 1 << c["reply"], c["reply"], 2, 1, sc[c["first"]], c["depth"], sc)
        assert (v, lv) == (c["value"], c["leaves"])
    for c in G["opening2"][:6]:
        v, lv = oracle.ab_subtree(4, 1 << c["first"], 0, c["first"], 1, -1, 0, c["depth"], sc)
        assert (v, lv) == (c["value"], c["leaves"])


def test_playout_exact(oracle):
    for c in G["playout_exact"]:
        p, plies = oracle.playout_exact(c["x"], c["o"], c["to_move"])
        assert all(abs(a - b) < 1e-12 for a, b in zip(p, c["p"])) and abs(plies - c["mean_plies"]) < 1e-10


def test_baseline_size_vectors_are_the_oracles(oracle):
    """the BASELINE-sized vectors the GPU tests and bench.py compare with (configs[3] eight-mark positions, configs[4]
    opening3 -d 10) were made by this oracle: the cheapest few are searched again here"""
    sc = oracle.tables()["scores_ab"]
    c3 = json.load(open(os.path.join(HERE, "golden", "c3_eight_mark_depth10.json")))["cases"]
    c = min(c3, key=lambda c: c["leaves"])
    vals, bm, bv, lv = oracle.ab1_root(c["x"], c["o"], 10, sc)
    assert (vals, bm, bv, lv) == (c["values"], c["best_mask"], c["best_value"], c["leaves"])
    so = oracle.tables()["scores_opening"]
    c4 = json.load(open(os.path.join(HERE, "golden", "c4_opening3_depth10.json")))
    assert [(c["first"], c["reply"]) for c in c4["cases"]] == oracle.opening3_subtrees()
    for c in sorted(c4["cases"], key=lambda c: c["evals"])[:6]:
        v, lv = oracle.ab_subtree(3, 1 << c["first"], 1 << c["reply"], c["reply"], 2, 1, so[c["first"]], 10, so)
        assert (v, lv, oracle.last_evals()) == (c["value"], c["leaves"], c["evals"])
    assert c4["total_evals"] == sum(c["evals"] for c in c4["cases"])
