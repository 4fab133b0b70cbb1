"""src/abbook (playoff5's "B" player) against the reference's own code.

tests/golden/ref_go_abbook_vectors.json was produced by executing src/abbook/abbook.go through the
in-memory Go->Python translation (oracle/tools/make_ref_golden_abbook.py):
  search  root values / chosen move / max / leaf count of ChooseMove with the book switched off
  depth   SetDepth for every move counter
  book    ChooseMove/MakeMove sequences against scripted opponents while the book is in progress
CPU part: the oracle (C and Python) reproduces `search`; the C++ host player reproduces the book
moves.  GPU part: dialect SQ_AB_ABBOOK reproduces `search`, and the host player the whole scripts."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import pyoracle as P

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
V = json.load(open(os.path.join(HERE, "golden", "ref_go_abbook_vectors.json")))
TABLES = json.load(open(os.path.join(HERE, "golden", "ref_tables.json")))
NO_MOVE = -(2 ** 31)


def _row(values):
    row = [NO_MOVE] * 25
    for k, v in values.items():
        row[int(k)] = v
    return row


def test_abbook_tables_are_src_alphabetas():
    """same line tables and default bias table as src/alphabeta: the dialects can share them"""
    assert V["scores"] == TABLES["ab"]["scores"]
    assert V["quads"] == TABLES["ab"]["quads"] and V["triplets"] == TABLES["ab"]["triplets"]


def test_oracle_matches_go_abbook_search(oracle):
    n = 0
    for c in V["search"]:
        vals, bm, bv, leaves = oracle.ab6_root(c["x"], c["o"], c["depth"], V["scores"])
        assert vals == _row(c["values"]), c
        assert (bv, leaves) == (c["max"], c["leaves"])
        assert c["move"] == min(k for k in range(25) if bm >> k & 1)
        if c["depth"] <= 4 or bin(c["x"] | c["o"]).count("1") >= 18:
            bd = [1 if c["x"] >> k & 1 else -1 if c["o"] >> k & 1 else 0 for k in range(25)]
            assert P.ab1_root(bd, c["depth"], V["scores"], cumulative=True) == vals
        n += 1
    assert n >= 40


@pytest.fixture(scope="module")
def host_lib():
    so = os.path.join(ROOT, "squava_b200", "host", "libsquava_host.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.dirname(so), "-s"])
    lib = C.CDLL(so)
    lib.sqh_abbook_script.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


def run_script(lib, case, book_only):
    script = (C.c_int * len(case["script"]))(*case["script"])
    moves = (C.c_int * 8)(); values = (C.c_int * 8)(); book = (C.c_int * 8)(); board = (C.c_int * 25)()
    still = C.c_int()
    n = lib.sqh_abbook_script(0 if case["role"] == "first" else 1, case["pick"], 2, script, len(case["script"]),
                              int(book_only), moves, values, book, 8, board, C.byref(still))
    return [(moves[i], values[i], bool(book[i])) for i in range(n)], list(board), bool(still.value)


def test_cpp_book_moves_match_go_abbook(host_lib):
    """the book state machine alone (no search, no GPU): every move made while bookInProgress"""
    n = 0
    for case in V["book"]:
        if any(m.get("book") and m["leaves"] for m in case["moves"]):
            continue                              # the book gave up and searched inside one call
        want = []
        for m in case["moves"]:
            if not m["book"]:
                break
            want.append((m["move"], m["value"], True))
        got, _, _ = run_script(host_lib, case, True)
        assert got == want, case
        n += len(want)
    assert n > 60


@pytest.mark.gpu
def test_gpu_matches_go_abbook_search(sq):
    sq.init()
    try:
        sc = np.array(V["scores"], np.int8)
        by_depth = {}
        for c in V["search"]:
            by_depth.setdefault(c["depth"], []).append(c)
        for d, cs in by_depth.items():
            b = np.array([(c["x"], c["o"]) for c in cs], dtype=np.uint32).view(sq.BOARD).reshape(-1)
            values, bv, bm, _ = sq.alphabeta_root(b, sq.AB_ABBOOK, d, sc)
            for i, c in enumerate(cs):
                assert values[i].tolist() == _row(c["values"]), (d, i)
                assert int(bv[i]) == c["max"] and c["move"] == min(k for k in range(25) if int(bm[i]) >> k & 1)
    finally:
        sq.shutdown()


@pytest.mark.gpu
def test_gpu_host_player_replays_go_abbook_scripts(sq, host_lib):
    """book moves, then the searched moves (depth 2, deterministic) and the final board"""
    sq.init()
    try:
        for case in V["book"]:
            got, board, still = run_script(host_lib, case, False)
            assert [g[0] for g in got] == [m["move"] for m in case["moves"]], case
            assert [g[1] for g in got] == [m["value"] for m in case["moves"]]
            assert board == case["board"] and still == case["still_in_book"]
    finally:
        sq.shutdown()
