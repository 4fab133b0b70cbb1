#!/usr/bin/env python3
"""Golden vectors from the reference's src/negascout/negascout.go (playoff5's "N" player), executed
through the in-memory Go->Python translation of go2py.py (no Go toolchain here; nothing of the
reference is copied into the repo).  Output: tests/golden/ref_go_negascout_vectors.json
  search   (*NegaScout).ChooseMove (negascout.go:81-122) under -D: every (cell, alpha) handed to
           movekeeper.SetMove in visiting order, the move returned, its value, leafNodeCount
  orders   reorderMoves (:411-499): orderedMoves[0] and [2] for the same positions
  static   staticValue (:168-230) on random boards, lines and all
  depth    SetDepth (:69-79) for moveCounter 0..24

TEST INFRASTRUCTURE; run once in the build container.
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
OUT = os.path.join(ROOT, "tests", "golden", "ref_go_negascout_vectors.json")


class Obj:
    pass


def load():
    mk = {"rand": Obj()}
    exec(compile(go2py.translate(open(os.path.join(REF, "src/movekeeper/movekeeper.go")).read(),
                                 {"MoveKeeper", "New", "MoveKeeper_SetMove", "MoveKeeper_ChooseMove"}),
                 "movekeeper.go (translated in memory)", "exec"), mk)
    recorded = []

    def new_keeper(mx, det):
        r = mk["New"](mx, det)

        def set_move(a, b, value):
            recorded.append((a, b, value))
            mk["MoveKeeper_SetMove"](r, a, b, value)
        r.SetMove = set_move
        r.ChooseMove = functools.partial(mk["MoveKeeper_ChooseMove"], r)
        return r
    movekeeper = Obj(); movekeeper.New = new_keeper
    said = []
    fmt = Obj(); fmt.Printf = lambda *a: said.append(a)
    src = open(os.path.join(REF, "src/negascout/negascout.go")).read()
    want = {"NegaScout", "indexedLosingTriplets", "indexedWinningQuads", "indexedCalcs", "calculateIndexedMatrices", "New",
            "NegaScout_MakeMove", "NegaScout_SetDepth", "NegaScout_ChooseMove", "NegaScout_staticValue", "NegaScout_negaScout",
            "NegaScout_reorderMoves", "NegaScout_FindWinner", "deadlyQuads", "checkableCells", "losingTriplets", "winningQuads",
            "scores", "initialOrderedMoves", "orderedMoves"}
    py = go2py.translate(src, want, {"New": ("indexedCalcs",)})
    ns = {"movekeeper": movekeeper, "fmt": fmt, "rand": Obj(),
          "new": lambda t: t(), "board": lambda: [[0] * 5 for _ in range(5)]}
    exec(compile(py, "negascout.go (translated in memory)", "exec"), ns)
    go2py.attach_methods(ns)
    # SetScores(false), negascout.go:380-386
    ns["scores"] = [[3, 3, 0, 3, 3], [3, 4, 1, 4, 3], [0, 1, 0, 1, 0], [3, 4, 1, 4, 3], [3, 3, 0, 3, 3]]
    return ns, recorded, said


def grid(x, o):
    return [[1 if (x >> (5 * i + j)) & 1 else -1 if (o >> (5 * i + j)) & 1 else 0 for j in range(5)] for i in range(5)]


def main():
    from squava_b200 import positions
    ns, recorded, said = load()
    out = {"source": "src/negascout/negascout.go of the reference, translated in memory by oracle/tools/go2py.py and "
                     "executed by oracle/tools/make_ref_golden_negascout.py"}
    ns["New"](True, 6)
    out["scores"] = [int(v) for row in ns["scores"] for v in row]
    out["initial_order"] = [5 * a + b for a, b in ns["initialOrderedMoves"]]
    out["deadly"] = [[5 * a + b for a, b in q] for q in ns["deadlyQuads"]]
    out["checkable"] = [5 * a + b for a, b in ns["checkableCells"]]
    depth = []
    for mc in range(25):
        q = ns["New"](True, 99)
        q.SetDepth(mc)
        depth.append(int(q.maxDepth))
    out["depth"] = depth

    g = np.random.Generator(np.random.PCG64(909))
    cases = []
    for marks, d, count in [(16, 8, 4), (14, 6, 4), (12, 6, 3), (12, 4, 4), (10, 4, 4), (10, 5, 2), (8, 4, 2), (8, 3, 3), (18, 10, 3),
                            (20, 10, 4), (22, 10, 3), (24, 8, 2), (6, 3, 2), (4, 2, 2), (2, 2, 2), (0, 1, 1), (14, 0, 2), (14, 1, 2),
                            (14, 2, 2), (14, 3, 2)]:
        for _ in range(count):
            x, o = positions.random_position(g, marks)
            if (x | o) == 0x1FFFFFF:
                continue
            q = ns["New"](True, d)
            q.bd = grid(x, o)
            del recorded[:]
            a, b, v, leaves = q.ChooseMove()
            cases.append({"x": x, "o": o, "depth": d, "set_moves": [[5 * i + j, int(val)] for i, j, val in recorded],
                          "move": 5 * a + b, "max": int(v), "leaves": int(leaves),
                          "order_min": [5 * i + j for i, j in ns["orderedMoves"][0]],
                          "order_max": [5 * i + j for i, j in ns["orderedMoves"][2]]})
        print("search", marks, d, flush=True)
    out["search"] = cases
    assert not said, said

    rng = random.Random(5)
    static = []
    for _ in range(600):
        fill = rng.random()
        cells = [rng.choice((1, -1)) if rng.random() < fill else 0 for _ in range(25)]
        q = ns["New"](True, 4)
        q.bd = [cells[5 * i:5 * i + 5] for i in range(5)]
        for ply in (3, 5):
            stop, val = q.staticValue(ply)
            static.append({"x": sum(1 << c for c in range(25) if cells[c] == 1), "o": sum(1 << c for c in range(25) if cells[c] == -1),
                           "ply": ply, "stop": bool(stop), "value": int(val)})
    out["static"] = static
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
