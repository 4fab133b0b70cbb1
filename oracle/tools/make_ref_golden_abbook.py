#!/usr/bin/env python3
"""Golden vectors from the reference's src/abbook/abbook.go (playoff5's "B" player), executed
through the in-memory Go->Python translation of go2py.py (no Go toolchain in this image; nothing of
the reference is copied into the repo).  Output: tests/golden/ref_go_abbook_vectors.json
  search   (*AlphaBetaBook).ChooseMove with the book switched off (abbook.go:104-130): the value of
           every root move as handed to movekeeper.SetMove, the move chosen under -D, max, leaf count
  depth    SetDepth (abbook.go:80-90) for moveCounter 0..24
  book     whole ChooseMove/MakeMove sequences while the book is in progress, as first and as second
           player, against scripted opponent replies, with rand.Intn pinned to each of 0..3

TEST INFRASTRUCTURE; run once in the build container.
"""
import functools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import go2py  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_go_abbook_vectors.json")


class Obj:
    pass


class Exit(Exception):
    pass


def load():
    mk_src = open(os.path.join(REF, "src/movekeeper/movekeeper.go")).read()
    mk = {"rand": Obj()}
    exec(compile(go2py.translate(mk_src, {"MoveKeeper", "New", "MoveKeeper_SetMove", "MoveKeeper_ChooseMove"}),
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
    rnd = Obj(); rnd.pick = 0
    rnd.Intn = lambda n: rnd.pick % n
    fmt = Obj(); fmt.Printf = lambda *a: None

    def _exit(code):
        raise Exit(code)
    os_ = Obj(); os_.Exit = _exit
    src = open(os.path.join(REF, "src/abbook/abbook.go")).read()
    want = {"AlphaBetaBook", "board", "indexedLosingTriplets", "indexedWinningQuads", "indexedCalcs", "calculateIndexedMatrices",
            "New", "AlphaBetaBook_MakeMove", "AlphaBetaBook_SetDepth", "AlphaBetaBook_ChooseMove", "AlphaBetaBook_deltaValue",
            "AlphaBetaBook_alphaBeta", "AlphaBetaBook_FindWinner", "AlphaBetaBook_bookDefend", "AlphaBetaBook_bookStart",
            "AlphaBetaBook_PrintBoard", "losingTriplets", "winningQuads", "scores", "firstMoves",
            "checkableCells"}
    py = go2py.translate(src, want, {"New": ("indexedCalcs",)})
    ns = {"movekeeper": movekeeper, "rand": rnd, "fmt": fmt, "os": os_,
          "new": lambda t: t(), "board": lambda: [[0] * 5 for _ in range(5)]}   # `type board [5][5]int`, new(board)
    exec(compile(py, "abbook.go (translated in memory)", "exec"), ns)
    go2py.attach_methods(ns)
    # SetScores(false), abbook.go:390-396 (a multi-line composite literal the translator does not take)
    ns["scores"] = [[3, 3, 0, 3, 3], [3, 4, 1, 4, 3], [0, 1, 0, 1, 0], [3, 4, 1, 4, 3], [3, 3, 0, 3, 3]]
    return ns, recorded, rnd


def squavathr_book():
    """squavathr.go's -B paths: bookStart (:735-810) and bookDefend (:662-733), with readMove scripted and
    printBoard silenced; records the "My move" lines, the returned move count and the board."""
    src = open(os.path.join(REF, "squavathr.go")).read()
    rnd = Obj(); rnd.pick = 0
    rnd.Intn = lambda n: rnd.pick % n
    said = []
    fmt = Obj()
    fmt.Printf = lambda f, *a: said.append([int(v) for v in a]) if f.startswith("My move") else None
    script = []

    def read_move(bd, printit):                # readMove, squavathr.go:559-610: occupied cells are refused
        while True:
            if not script:
                raise Exit(0)                  # EOF on stdin exits the program (:566-568)
            c = script.pop(0)
            if bd[c // 5][c % 5] == 0:
                return c // 5, c % 5

    def _exit(code):
        raise Exit(code)
    os_ = Obj(); os_.Exit = _exit
    ns = {"rand": rnd, "fmt": fmt, "os": os_, "readMove": read_move, "printBoard": lambda bd: None}
    exec(compile(go2py.translate(src, {"bookDefend", "bookStart", "firstMoves"}), "squavathr.go (translated in memory)", "exec"), ns)
    out = []
    for pick in range(4):
        for replies in ([12, 7, 3], [24, 13, 2], [18, 3, 9], [19, 9, 14], [23, 4, 11], [6, 20, 1], [21, 8, 13], [16, 15, 22], [18]):
            rnd.pick = pick
            del said[:]
            script[:] = list(replies)
            bd = [[0] * 5 for _ in range(5)]
            try:
                count = int(ns["bookStart"](bd))
            except Exit as e:
                count = -1 - int(e.args[0])
            out.append({"role": "start", "pick": pick, "replies": replies, "said": [list(v) for v in said], "count": count,
                        "unread": len(script), "board": [int(v) for row in bd for v in row]})
    for first, replies in ((0, [18, 7]), (6, [24, 1]), (12, [0, 3]), (4, [16, 9]), (21, [3, 14]), (1, [19, 5]), (24, [6, 2]),
                           (9, [15, 8]), (3, [18, 6]), (18, [0, 21]), (7, [13])):
        del said[:]
        script[:] = list(replies)
        bd = [[0] * 5 for _ in range(5)]
        bd[first // 5][first % 5] = ns["MINIMIZER"]                               # main(), squavathr.go:97-99
        try:
            count = int(ns["bookDefend"](bd, first // 5, first % 5))
        except Exit as e:
            count = -1 - int(e.args[0])
        out.append({"role": "defend", "first": first, "replies": replies, "said": [list(v) for v in said], "count": count,
                    "unread": len(script), "board": [int(v) for row in bd for v in row]})
    return out


def grid(x, o):
    return [[1 if (x >> (5 * i + j)) & 1 else -1 if (o >> (5 * i + j)) & 1 else 0 for j in range(5)] for i in range(5)]


def main():
    from squava_b200 import positions
    ns, recorded, rnd = load()
    out = {"source": "src/abbook/abbook.go of the reference, translated in memory by oracle/tools/go2py.py and executed "
                     "by oracle/tools/make_ref_golden_abbook.py"}
    ns["New"](True, 6)
    out["scores"] = [int(v) for row in ns["scores"] for v in row]
    out["quads"] = [[5 * a + b for a, b in q] for q in ns["winningQuads"]]
    out["triplets"] = [[5 * a + b for a, b in t] for t in ns["losingTriplets"]]

    depth = []
    for mc in range(25):
        q = ns["New"](True, 99)
        q.SetDepth(mc)
        depth.append(int(q.maxDepth))
    out["depth"] = depth

    g = np.random.Generator(np.random.PCG64(777))
    cases = []
    for marks, d, count in [(14, 10, 4), (12, 8, 5), (10, 8, 2), (10, 6, 4), (8, 6, 2), (16, 10, 4), (18, 10, 3), (20, 10, 3),
                            (22, 8, 2), (24, 6, 2), (6, 4, 3), (2, 4, 2), (0, 3, 1), (14, 0, 2), (14, 1, 2)]:
        for _ in range(count):
            x, o = positions.random_position(g, marks)
            if (x | o) == 0x1FFFFFF:
                continue
            q = ns["New"](True, d)
            q.bd = grid(x, o)
            q.bookInProgress = False
            del recorded[:]
            a, b, v, leaves = q.ChooseMove()
            cases.append({"x": x, "o": o, "depth": d, "values": {str(5 * i + j): int(val) for i, j, val in recorded},
                          "move": 5 * a + b, "max": int(v), "leaves": int(leaves)})
        print("search", marks, d, flush=True)
    out["search"] = cases

    # the book: scripted opponents.  Each script is the list of opponent replies (cells), consumed whenever it
    # is the opponent's turn; the engine's ChooseMove is recorded while bookInProgress was set when it was called.
    book = []
    scripts_first = [[12, 7], [24, 13], [18, 3], [19, 9], [23, 4], [6, 20], [21, 8], [16, 15]]
    for pick in range(4):
        for script in scripts_first:
            rnd.pick = pick
            q = ns["New"](True, 2)
            moves, ok = [], True
            for reply in script + [None]:
                in_book = bool(q.bookInProgress)
                try:
                    a, b, v, leaves = q.ChooseMove()
                except Exit as e:
                    moves.append({"exit": int(e.args[0])}); ok = False
                    break
                moves.append({"move": 5 * a + b, "value": int(v), "leaves": int(leaves), "book": in_book})
                if reply is None or q.bd[reply // 5][reply % 5] != 0:
                    break
                q.MakeMove(reply // 5, reply % 5, ns["MINIMIZER"])
            book.append({"role": "first", "pick": pick, "script": script, "moves": moves,
                         "board": [int(v) for row in q.bd for v in row], "still_in_book": bool(q.bookInProgress)})
    scripts_second = [[0, 18], [6, 24], [12, 0], [4, 16], [21, 3], [1, 19], [24, 6], [9, 15], [3, 18]]
    for script in scripts_second:
        q = ns["New"](True, 2)
        moves = []
        for reply in script:
            if q.bd[reply // 5][reply % 5] != 0:
                break
            q.MakeMove(reply // 5, reply % 5, ns["MINIMIZER"])
            in_book = bool(q.bookInProgress)
            a, b, v, leaves = q.ChooseMove()
            moves.append({"move": 5 * a + b, "value": int(v), "leaves": int(leaves), "book": in_book})
        book.append({"role": "second", "pick": 0, "script": script, "moves": moves,
                     "board": [int(v) for row in q.bd for v in row], "still_in_book": bool(q.bookInProgress)})
    out["book"] = book
    out["squavathr_book"] = squavathr_book()
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
