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
