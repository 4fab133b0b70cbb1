#!/usr/bin/env python3
"""Parse the line/scan-order tables out of the reference's Go sources and write
them to tests/golden/ref_tables.json.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (where /root/reference is
mounted read-only); the JSON it writes is committed so the same comparison runs
where the reference is absent (the GPU box).  Nothing here is imported by the
product.

What is extracted (all cell coordinates flattened to c = 5*x + y):
  mcts.important_cells / mcts.quads / mcts.triplets   src/mcts/mcts.go:404-446
  ab.quads / ab.triplets (flat, order matters)        src/alphabeta/alphabeta.go:249-328
  ab.scores (default bias)                            src/alphabeta/alphabeta.go:343-349
  thr.no2 / thr.no_middle2                            squavathr.go:297-311
  op3.ordered_cells / op3.scores / op3.matrices / op3.first_moves
                                                      opening3.go:328-372,379-463
  op2.ordered_cells / op2.scores / op2.unique_cells   opening2.go:262-313
and, as a consistency check, that the copies in squavam.go / squavathr.go /
opening2.go / opening3.go are identical to the package copies.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(
    os.path.dirname(__file__), "..", "..", "tests", "golden", "ref_tables.json")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def grab_literal(src, name):
    """Return the text of the composite literal assigned to `var name ... = T{ ... }`."""
    m = re.search(r"var\s+" + re.escape(name) + r"\b[^=\n]*=\s*[^{\n]*\{", src)
    if not m:
        raise KeyError(name)
    i = m.end() - 1
    depth = 0
    for j in range(i, len(src)):
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i:j + 1]
    raise ValueError(name)


def to_nested(lit):
    """Go composite literal of ints -> nested python lists (type prefixes dropped)."""
    lit = re.sub(r"(\[\d*\])+int", "", lit)       # [][]int, [2]int, [5][2]int ...
    lit = lit.replace("{", "[").replace("}", "]")
    lit = re.sub(r",\s*\]", "]", lit)
    lit = re.sub(r"(-?\d+)\s*-\s*(\d+)", lambda m: str(int(m.group(1)) - int(m.group(2))), lit)
    return json.loads(lit)


def load(fn):
    with open(os.path.join(REF, fn)) as f:
        return strip_comments(f.read())


def flat(pairs):
    return [5 * p[0] + p[1] for p in pairs]


def main():
    out = {"source": "bediger4000/squava (parsed by oracle/tools/extract_ref_tables.py)"}

    mcts = load("src/mcts/mcts.go")
    imp = to_nested(grab_literal(mcts, "importantCells"))
    q1 = to_nested(grab_literal(mcts, "winningQuads"))
    t1 = to_nested(grab_literal(mcts, "losingTriplets"))
    out["mcts"] = {"important_cells": imp, "quads": q1, "triplets": t1}
    sm = load("squavam.go")
    assert to_nested(grab_literal(sm, "importantCells")) == imp
    assert to_nested(grab_literal(sm, "winningQuads")) == q1
    assert to_nested(grab_literal(sm, "losingTriplets")) == t1

    ab = load("src/alphabeta/alphabeta.go")
    q2 = [flat(q) for q in to_nested(grab_literal(ab, "winningQuads"))]
    t2 = [flat(t) for t in to_nested(grab_literal(ab, "losingTriplets"))]
    m = re.search(r"scores = \[5\]\[5\]int(\{.*?\n\t\t\})", ab, flags=re.S)
    ab_scores = [v for row in to_nested(m.group(1)) for v in row]
    out["ab"] = {"quads": q2, "triplets": t2, "scores": ab_scores,
                 "no2": [flat(t) for t in to_nested(grab_literal(ab, "no2"))],
                 "no_middle2": [flat(q) for q in to_nested(grab_literal(ab, "noMiddle2"))]}

    thr = load("squavathr.go")
    assert [flat(q) for q in to_nested(grab_literal(thr, "winningQuads"))] == q2
    assert [flat(t) for t in to_nested(grab_literal(thr, "losingTriplets"))] == t2
    m = re.search(r"scores = \[5\]\[5\]int(\{.*?\n\t\t\})", thr, flags=re.S)
    out["thr"] = {"scores": [v for row in to_nested(m.group(1)) for v in row],
                  "no2": [flat(t) for t in to_nested(grab_literal(thr, "no2"))],
                  "no_middle2": [flat(q) for q in to_nested(grab_literal(thr, "noMiddle2"))]}

    for key, fn in (("op2", "opening2.go"), ("op3", "opening3.go")):
        src = load(fn)
        assert [flat(q) for q in to_nested(grab_literal(src, "winningQuads"))] == q2
        assert [flat(t) for t in to_nested(grab_literal(src, "losingTriplets"))] == t2
        d = {"ordered_cells": flat(to_nested(grab_literal(src, "orderedCells"))),
             "scores": [v for row in to_nested(grab_literal(src, "scores")) for v in row],
             "unique_cells": flat(to_nested(grab_literal(src, "uniqueCells")))}
        if key == "op3":
            mats = to_nested(grab_literal(src, "Matrices"))
            d["matrices"] = [[5 * mt[x][y][0] + mt[x][y][1] for x in range(5) for y in range(5)]
                             for mt in mats]
            d["first_moves"] = flat(to_nested(grab_literal(src, "firstMoves")))
        out[key] = d

    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
