#!/usr/bin/env python3
"""Golden vectors from the reference's own Go code, executed through a mechanical Go->Python
translation (oracle/tools/go2py.py) because this image has no Go toolchain.  The Go files are read
from /root/reference at run time, the needed declarations are translated in memory and run; nothing
of the reference is copied into this repo.  What cannot be translated mechanically (goroutines and
channels in squavathr's chooseMove, flag/time/os in the opening programs' main) is re-driven here
line for line and cites the lines.

Output: tests/golden/ref_go_vectors.json
  alphabeta   src/alphabeta ChooseMove (alphabeta.go:92-117) with src/movekeeper: per position the
              value of every root move (as handed to movekeeper.SetMove), the chosen move under -D,
              the maximum and leafNodeCount; with deltaValue and with deltaValue2 (SetAvoid)
  squavathr   squavathr.go chooseMove (:236-269) + worker (:448-466): root values and leaf counts
  opening3    opening3.go main loop (:66-128): the 85 (first, reply) pairs with val and leafNodes
  opening2    opening2.go main loop (:62-71)
  winners     (*AlphaBeta).FindWinner (alphabeta.go:355-378), (*MCTS).FindWinner / GetMoves /
              GetResult (mcts.go:92-121,310-388) on random boards
"""
import functools
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import go2py  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_go_vectors.json")


class Obj:
    pass


def load(path, want, globals_written=None, extra=None):
    src = open(os.path.join(REF, path)).read()
    py = go2py.translate(src, set(want), globals_written)
    ns = dict(extra or {})
    exec(compile(py, path + " (translated in memory)", "exec"), ns)
    return ns


def grid(x, o):
    return [[1 if (x >> (5 * i + j)) & 1 else -1 if (o >> (5 * i + j)) & 1 else 0 for j in range(5)] for i in range(5)]


# ---- src/alphabeta + src/movekeeper ---------------------------------------------------------
def make_alphabeta():
    mk = load("src/movekeeper/movekeeper.go", ["MoveKeeper", "New", "MoveKeeper_SetMove", "MoveKeeper_ChooseMove"],
              extra={"rand": Obj()})
    recorded = []

    def new_keeper(mx, det):
        r = mk["New"](mx, det)

        def set_move(a, b, value):
            recorded.append((a, b, value))
            mk["MoveKeeper_SetMove"](r, a, b, value)
        r.SetMove = set_move
        r.ChooseMove = functools.partial(mk["MoveKeeper_ChooseMove"], r)
        return r
    movekeeper = Obj()
    movekeeper.New = new_keeper
    ab = load("src/alphabeta/alphabeta.go",
              ["indexedLosingTriplets", "indexedWinningQuads", "calculateIndexedMatrices", "deltaValue", "deltaValue2",
               "AlphaBeta_alphaBeta", "AlphaBeta_ChooseMove", "AlphaBeta_FindWinner", "AlphaBeta_MakeMove",
               "losingTriplets", "winningQuads", "scores", "noMiddle2", "no2"], extra={"movekeeper": movekeeper})
    ab["calculateIndexedMatrices"]()                         # alphabeta.go:53-56

    def player(x, o, depth, avoid, scores):
        p = Obj()
        p.bd = grid(x, o)
        p.maxDepth, p.deterministic, p.leafNodeCount = depth, True, 0
        p.boardValue = ab["deltaValue2"] if avoid else ab["deltaValue"]          # alphabeta.go:62,473-475
        p.alphaBeta = functools.partial(ab["AlphaBeta_alphaBeta"], p)
        p.MakeMove = functools.partial(ab["AlphaBeta_MakeMove"], p)
        ab["scores"] = [list(scores[5 * i:5 * i + 5]) for i in range(5)]        # SetScores(false), :343-349
        return p

    def choose(x, o, depth, avoid, scores):
        del recorded[:]
        p = player(x, o, depth, avoid, scores)
        if all(v != 0 for row in p.bd for v in row):
            return None
        a, b, v, leaves = ab["AlphaBeta_ChooseMove"](p)
        return {"values": {str(5 * i + j): int(val) for i, j, val in recorded}, "move": 5 * a + b, "max": int(v),
                "leaves": int(leaves)}

    def find_winner(x, o):
        p = Obj(); p.bd = grid(x, o)
        return int(ab["AlphaBeta_FindWinner"](p))
    return choose, find_winner


# ---- squavathr.go -------------------------------------------------------------------------------
def make_squavathr():
    ns = load("squavathr.go", ["indexedLosingTriplets", "indexedWinningQuads", "deltaValue", "alphaBeta", "losingTriplets",
                               "winningQuads", "scores", "noMiddle2", "no2", "findWinner"])
    for triplet in ns["losingTriplets"]:                      # main(), squavathr.go:61-71
        for pair in triplet:
            ns["indexedLosingTriplets"][pair[0]][pair[1]].append(triplet)
    for quad in ns["winningQuads"]:
        for pair in quad:
            ns["indexedWinningQuads"][pair[0]][pair[1]].append(quad)

    def choose(x, o, max_depth, scores):
        ns["scores"] = [list(scores[5 * i:5 * i + 5]) for i in range(5)]         # setScores(false), :621-629
        bd = grid(x, o)
        values, leaf_nodes = {}, 0
        max_depth -= 1                                                           # :239
        for i in range(5):                                                       # :241-253
            for j in range(5):
                if bd[i][j] == ns["UNSET"]:
                    stop, val = ns["deltaValue"](max_depth, bd, 0, i, j, 0)
                    if stop:
                        values[5 * i + j] = int(val)
                        leaf_nodes += 1
                    else:
                        gs = [row[:] for row in bd]                              # newState, :210-233
                        gs[i][j] = ns["MAXIMIZER"]
                        v, lv = ns["alphaBeta"](max_depth, gs, 1, ns["MINIMIZER"], 2 * ns["LOSS"], 2 * ns["WIN"], i, j, val)  # worker :451-460
                        values[5 * i + j] = int(v)
                        leaf_nodes += lv
        return {"values": {str(k): v for k, v in values.items()}, "leaves": int(leaf_nodes)}
    return choose


# ---- opening2.go / opening3.go ---------------------------------------------------------------------
def make_opening(which):
    fn = "opening%d.go" % which
    ns = load(fn, ["indexedLosingTriplets", "indexedWinningQuads", "alphaBeta", "losingTriplets", "winningQuads", "scores",
                   "orderedCells", "Matrices", "firstMoves", "uniqueCells", "maxDepth", "leafNodes"],
              globals_written={"alphaBeta": ("leafNodes",)})
    for triplet in ns["losingTriplets"]:
        for pair in triplet:
            ns["indexedLosingTriplets"][pair[0]][pair[1]].append(triplet)
    for quad in ns["winningQuads"]:
        for pair in quad:
            ns["indexedWinningQuads"][pair[0]][pair[1]].append(quad)
    return ns


def run_opening2(ns, depth):
    ns["maxDepth"] = depth
    out = []
    bd = [[0] * 5 for _ in range(5)]
    for cell in ns["uniqueCells"]:                                             # opening2.go:62-71
        i, j = cell
        bd[i][j] = ns["MAXIMIZER"]
        ns["leafNodes"] = 0
        val = ns["alphaBeta"](bd, 1, ns["MINIMIZER"], ns["LOSS"], ns["WIN"], i, j, 0)
        bd[i][j] = ns["UNSET"]
        out.append({"first": 5 * i + j, "value": int(val), "leaves": int(ns["leafNodes"])})
    return out


def run_opening3(ns, depth):
    ns["maxDepth"] = depth
    M, first_moves = ns["Matrices"], ns["firstMoves"]
    computed = [[[0] * 5 for _ in range(5)] for _ in range(6)]
    for k, cell in enumerate(first_moves):
        computed[k][cell[0]][cell[1]] = ns["MAXIMIZER"]
    bd = [[0] * 5 for _ in range(5)]
    out = []
    for fi, cell in enumerate(first_moves):                                    # opening3.go:66-128
        x, y = cell
        for p in range(5):
            for q in range(5):
                if p != x or q != y:
                    matched = False
                    for m in M:
                        a, b = m[x][y]
                        c, d = m[p][q]
                        for board in computed:
                            if board[a][b] == ns["MAXIMIZER"] and board[c][d] == ns["MINIMIZER"]:
                                matched = True
                                break
                        if matched:
                            break
                    if not matched:
                        computed[fi][p][q] = ns["MINIMIZER"]
                        bd[x][y] = ns["MAXIMIZER"]
                        bd[p][q] = ns["MINIMIZER"]
                        delta = ns["scores"][x][y]
                        ns["leafNodes"] = 0
                        val = ns["alphaBeta"](bd, 2, ns["MAXIMIZER"], ns["LOSS"], ns["WIN"], p, q, delta)
                        bd[x][y] = ns["UNSET"]
                        bd[p][q] = ns["UNSET"]
                        out.append({"first": 5 * x + y, "reply": 5 * p + q, "value": int(val), "delta": int(delta),
                                    "leaves": int(ns["leafNodes"])})
    return out


# ---- src/mcts/mcts.go ---------------------------------------------------------------------------------
def make_mcts():
    ns = load("src/mcts/mcts.go", ["importantCells", "winningQuads", "losingTriplets", "MCTS_FindWinner",
                                   "GameState_GetMoves", "GameState_GetResult"])

    def describe(x, o):
        st = Obj()
        st.board = [1 if (x >> c) & 1 else -1 if (o >> c) & 1 else 0 for c in range(25)]
        st.cachedResults = [-1.0, -1.0, -1.0]                                   # resetCachedResults, mcts.go:286-290
        moves, term = ns["GameState_GetMoves"](st)
        p = Obj(); p.game = st
        rec = {"x": x, "o": o, "terminal": bool(term), "moves": list(moves), "winner": int(ns["MCTS_FindWinner"](p))}
        st.cachedResults = [-1.0, -1.0, -1.0]
        rec["result_x"] = float(ns["GameState_GetResult"](st, 1))
        st.cachedResults = [-1.0, -1.0, -1.0]
        rec["result_o"] = float(ns["GameState_GetResult"](st, -1))
        return rec
    return describe


def main():
    from squava_b200 import positions
    tables = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_tables.json")))
    sc_ab, sc_thr = tables["ab"]["scores"], tables["thr"]["scores"]
    g = np.random.Generator(np.random.PCG64(4242))
    out = {"source": "Go files of the reference, translated in memory by oracle/tools/go2py.py and executed by "
                     "oracle/tools/make_ref_golden_go.py"}

    choose, ab_winner = make_alphabeta()
    cases = []
    plan = [(14, 10, 6), (12, 10, 3), (16, 12, 5), (12, 8, 4), (10, 8, 2), (10, 6, 4), (8, 6, 2), (18, 12, 4), (20, 12, 3),
            (22, 10, 2), (24, 6, 2), (6, 4, 3), (2, 4, 2), (0, 3, 1), (14, 0, 2), (14, 1, 2)]
    for marks, depth, count in plan:
        for _ in range(count):
            x, o = positions.random_position(g, marks)
            for avoid in (False, True):
                r = choose(x, o, depth, avoid, sc_ab)
                if r:
                    r.update({"x": x, "o": o, "depth": depth, "avoid": avoid})
                    cases.append(r)
        print("alphabeta", marks, depth, flush=True)
    # Appendix B of SURVEY.md: two of its positions, small enough for the translated code
    for x, o in ((0x05c2408, 0x1800274), (0x0412825, 0x098c0c0)):
        r = choose(x, o, 10, False, sc_ab)
        r.update({"x": x, "o": o, "depth": 10, "avoid": False})
        cases.append(r)
    out["alphabeta"] = cases

    thr = make_squavathr()
    cases = []
    for marks, depth, count in [(14, 10, 4), (12, 9, 3), (16, 12, 4), (10, 7, 3), (18, 15, 3), (8, 5, 3), (20, 15, 2), (4, 3, 2), (14, 1, 2), (14, 2, 2)]:
        for _ in range(count):
            x, o = positions.random_position(g, marks)
            r = thr(x, o, depth, sc_thr)
            r.update({"x": x, "o": o, "depth": depth})
            cases.append(r)
        print("squavathr", marks, depth, flush=True)
    out["squavathr"] = cases

    o3 = make_opening(3)
    out["opening3"] = {str(d): run_opening3(o3, d) for d in (3, 4, 5)}
    print("opening3 done", flush=True)
    o2 = make_opening(2)
    out["opening2"] = {str(d): run_opening2(o2, d) for d in (2, 3, 4, 5, 6)}
    print("opening2 done", flush=True)

    describe = make_mcts()
    rng = random.Random(99)
    boards = []
    for _ in range(1500):
        fill = rng.random()
        cells = [rng.choice((1, -1)) if rng.random() < fill else 0 for _ in range(25)]
        x = sum(1 << c for c in range(25) if cells[c] == 1)
        o = sum(1 << c for c in range(25) if cells[c] == -1)
        rec = describe(x, o)
        rec["ab_winner"] = ab_winner(x, o)
        boards.append(rec)
    for _ in range(60):
        k = int(g.integers(5, 25))
        cellsp = g.permutation(25)[:k]
        x = sum(1 << int(c) for c in cellsp[0::2]); o = sum(1 << int(c) for c in cellsp[1::2])
        rec = describe(x, o)
        rec["ab_winner"] = ab_winner(x, o)
        boards.append(rec)
    out["winners"] = boards
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
