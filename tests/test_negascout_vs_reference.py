"""src/negascout (playoff5's "N" player) against the reference's own code.

tests/golden/ref_go_negascout_vectors.json was produced by executing src/negascout/negascout.go
through the in-memory Go->Python translation (oracle/tools/make_ref_golden_negascout.py).  The
search is order- and window-dependent, so the vectors hold the whole SetMove sequence and the leaf
counter.  CPU part: the oracle (C and Python).  GPU part: sq_negascout_root."""
import json
import os

import numpy as np
import pytest

import pyoracle as P

HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "ref_go_negascout_vectors.json")))
TABLES = json.load(open(os.path.join(HERE, "golden", "ref_tables.json")))
NO_MOVE = -(2 ** 31)


def test_negascout_constants():
    assert V["scores"] == TABLES["ab"]["scores"]
    assert V["initial_order"] == P.NS_INITIAL_ORDER and V["deadly"] == P.NS_DEADLY and V["checkable"] == P.IMPORTANT
    assert V["depth"] == [6] * 4 + [8] * 7 + [10] * 14          # SetDepth, negascout.go:69-79


def test_oracle_matches_go_negascout_search(oracle):
    n = 0
    for c in V["search"]:
        vals, order, bm, bv, fb, leaves = oracle.negascout_root(c["x"], c["o"], c["depth"], V["scores"])
        assert [[k, vals[k]] for k in order] == c["set_moves"], c
        assert (fb, bv, leaves) == (c["move"], c["max"], c["leaves"])
        bd = [1 if c["x"] >> k & 1 else -1 if c["o"] >> k & 1 else 0 for k in range(25)]
        assert list(P.ns_reorder(bd)) == [c["order_min"], c["order_max"]]
        if c["leaves"] < 30000:
            pv, po, pl = P.negascout_root(bd, c["depth"], V["scores"])
            assert (pv, po, pl) == (vals, order, leaves)
        n += 1
    assert n >= 50


def test_python_static_value_matches_go():
    for s in V["static"]:
        bd = [1 if s["x"] >> k & 1 else -1 if s["o"] >> k & 1 else 0 for k in range(25)]
        assert P.ns_static(bd, s["ply"], 4, V["scores"]) == (s["stop"], s["value"]), s


@pytest.mark.gpu
def test_gpu_matches_go_negascout(sq):
    sq.init()
    try:
        sc = np.array(V["scores"], np.int8)
        by_depth = {}
        for c in V["search"]:
            by_depth.setdefault(c["depth"], []).append(c)
        for d, cs in by_depth.items():
            b = np.array([(c["x"], c["o"]) for c in cs], dtype=np.uint32).view(sq.BOARD).reshape(-1)
            r = sq.negascout_root(b, d, sc)
            for i, c in enumerate(cs):
                order = [int(k) for k in r["visit_order"][i] if k >= 0]
                assert [[k, int(r["values"][i][k])] for k in order] == c["set_moves"], (d, i)
                assert (int(r["first_best"][i]), int(r["best_value"][i]), int(r["nodes"][i])) == (c["move"], c["max"], c["leaves"])
    finally:
        sq.shutdown()
