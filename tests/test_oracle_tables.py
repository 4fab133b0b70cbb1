"""The oracle's tables against tests/golden/ref_tables.json (parsed from the reference's
Go sources by oracle/tools/extract_ref_tables.py) and structural invariants."""
import itertools
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_tables.json")))


def all_lines(length):
    out = set()
    for x in range(5):
        for y in range(5):
            for dx, dy in ((1, 0), (0, 1), (1, 1), (-1, 1)):
                cells = [(x + i * dx, y + i * dy) for i in range(length)]
                if all(0 <= a < 5 and 0 <= b < 5 for a, b in cells):
                    out.add(frozenset(5 * a + b for a, b in cells))
    return out


def test_tables_match_reference(oracle):
    t = oracle.tables()
    assert t["important"] == GOLD["mcts"]["important_cells"]
    assert t["mcts_quads"] == GOLD["mcts"]["quads"]
    assert t["mcts_triplets"] == GOLD["mcts"]["triplets"]
    assert t["ab_quads"] == GOLD["ab"]["quads"]
    assert t["ab_triplets"] == GOLD["ab"]["triplets"]
    assert t["ordered"] == GOLD["op3"]["ordered_cells"] == GOLD["op2"]["ordered_cells"]
    assert t["scores_ab"] == GOLD["ab"]["scores"] == GOLD["thr"]["scores"]
    assert t["scores_opening"] == GOLD["op3"]["scores"] == GOLD["op2"]["scores"]
    assert t["no2"] == GOLD["thr"]["no2"] == GOLD["ab"]["no2"]
    assert t["no_middle2"] == GOLD["thr"]["no_middle2"] == GOLD["ab"]["no_middle2"]
    assert t["matrices"] == GOLD["op3"]["matrices"]
    assert t["first_moves"] == GOLD["op3"]["first_moves"] == GOLD["op3"]["unique_cells"]


def test_tables_cover_every_line_exactly():
    q28, t48 = all_lines(4), all_lines(3)
    assert len(q28) == 28 and len(t48) == 48
    assert {frozenset(q) for q in GOLD["ab"]["quads"]} == q28
    assert {frozenset(t) for t in GOLD["ab"]["triplets"]} == t48
    mq = [frozenset(q) for c in GOLD["mcts"]["quads"] for q in c]
    mt = [frozenset(t) for c in GOLD["mcts"]["triplets"] for t in c]
    assert len(mq) == 28 and set(mq) == q28          # each quad listed once
    assert len(mt) == 72 and set(mt) == t48          # 48 lines + 24 repeats
    # every line listed under an important cell contains that cell
    for c, (qs, ts) in enumerate(zip(GOLD["mcts"]["quads"], GOLD["mcts"]["triplets"])):
        assert all(c in q for q in qs) and all(c in t for t in ts)
        assert bool(qs or ts) == (c in GOLD["mcts"]["important_cells"])


def test_matrices_are_d4_and_map_lines_to_lines():
    mats = GOLD["op3"]["matrices"]
    assert len({tuple(m) for m in mats}) == 8
    q28, t48 = all_lines(4), all_lines(3)
    for m in mats:
        assert sorted(m) == list(range(25))
        assert {frozenset(m[c] for c in q) for q in q28} == q28
        assert {frozenset(m[c] for c in t) for t in t48} == t48
    # closed under composition
    S = {tuple(m) for m in mats}
    for a, b in itertools.product(mats, mats):
        assert tuple(a[b[c]] for c in range(25)) in S


def test_opening3_dedup_yields_85(oracle):
    import pyoracle
    c_list = oracle.opening3_subtrees()
    assert len(c_list) == 85
    assert c_list == pyoracle.opening3_subtrees()
    per_first = [sum(1 for f, _ in c_list if f == fm) for fm in GOLD["op3"]["first_moves"]]
    assert per_first == [14, 24, 14, 14, 14, 5]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_golden_is_fresh(tmp_path):
    out = tmp_path / "t.json"
    subprocess.check_call([sys.executable, os.path.join(HERE, "..", "oracle", "tools", "extract_ref_tables.py"),
                           "/root/reference", str(out)], stdout=subprocess.DEVNULL)
    assert json.load(open(out)) == GOLD
