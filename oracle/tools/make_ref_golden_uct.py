#!/usr/bin/env python3
"""Golden trees from the reference's own UCT driver: src/mcts/mcts.go (UCB1 with the factor 2) and
squavam.go (without it), translated in memory by go2py.py and executed with a DETERMINISTIC
rand.Intn (a SplitMix64 counter stream, value % n), so that another implementation fed the same
stream must grow the same tree.  Output: tests/golden/ref_go_uct_vectors.json -- per case the root's
children in creation order (move, visits, wins), the move UCT returns, int(1000*UCB1) and how many
draws were consumed.

TEST INFRASTRUCTURE; run once in the build container.
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, HERE)
import go2py  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_go_uct_vectors.json")
MASK = (1 << 64) - 1


def splitmix(k):
    z = (k * 0x9E3779B97F4A7C15) & MASK
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & MASK
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


class Obj:
    pass


WANT = ["GameState", "Node", "UCT", "NewNode", "Node_bestMove", "Node_UCTSelectChild", "Node_UCB1", "Node_AddChild",
        "Node_Update", "NewGameState", "GameState_Clone", "GameState_resetCachedResults", "GameState_DoMove",
        "GameState_GetMoves", "GameState_GetResult", "importantCells", "winningQuads", "losingTriplets"]


def load(path):
    src = open(os.path.join(REF, path)).read()
    py = go2py.translate(src, set(WANT))
    m = Obj()
    m.Sqrt = lambda v: math.sqrt(v) if v >= 0 else float("nan")
    m.Log = lambda v: math.log(v) if v > 0 else float("-inf")
    m.SmallestNonzeroFloat64 = 5e-324
    counter = [0]
    r = Obj()

    def intn(n):
        counter[0] += 1
        return splitmix(counter[0]) % n
    r.Intn = intn
    f = Obj(); f.Printf = lambda *a: None
    ns = {"math": m, "rand": r, "fmt": f}
    exec(compile(py, path + " (translated in memory)", "exec"), ns)
    go2py.attach_methods(ns)
    return ns, counter


def run(ns, counter, x, o, pjm, itermax, uctk, which):
    counter[0] = 0
    st = ns["NewGameState"]()
    st.board = [1 if (x >> c) & 1 else -1 if (o >> c) & 1 else 0 for c in range(25)]
    st.playerJustMoved = pjm
    res = ns["UCT"](st, itermax, uctk, None)
    if which == "mcts":
        best, leaves, value = res
    else:
        best, leaves = res, itermax
        value = int(1000. * best.UCB1(uctk))
    root = best.parentNode
    return {"which": which, "x": x, "o": o, "pjm": pjm, "itermax": itermax, "uctk": uctk,
            "children": [[c.move, c.visits, c.wins] for c in root.childNodes], "best": best.move, "value": value,
            "draws": counter[0], "root_visits": root.visits}


def main():
    cases = []
    for which, path in (("mcts", "src/mcts/mcts.go"), ("squavam", "squavam.go")):
        ns, counter = load(path)
        for (x, o, pjm) in ((0, 0, -1), (0, 0, 1), (0x1080038, 0x0c00181, -1), (0x0412825, 0x098c0c0, -1),
                            (1 << 12, 0, 1), (0x0504208, 0x00280a1, -1), (0x0c0090a, 0x02400d4, -1), (0x1408051, 0x0a21500 | (1 << 1), 1)):
            for itermax, uctk in ((30, 1.0), (400, 1.0), (3000, 0.5), (2500, 1.0)):
                cases.append(run(ns, counter, x, o, pjm, itermax, uctk, which))
        print(which, "done", flush=True)
    with open(OUT, "w") as f:
        json.dump({"source": "src/mcts/mcts.go and squavam.go of the reference, translated in memory (go2py.py) and "
                             "executed by oracle/tools/make_ref_golden_uct.py with rand.Intn(n) = splitmix64(k) % n",
                   "cases": cases}, f, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
