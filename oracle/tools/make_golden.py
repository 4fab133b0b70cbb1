#!/usr/bin/env python3
"""Write tests/golden/oracle_vectors.json: outputs of the C oracle on seeded positions.

TEST INFRASTRUCTURE.  These vectors come from THIS repo's oracle, not from the
reference (which cannot run here: no Go toolchain) -- they pin the oracle against
regressions and give the GPU tests fixed inputs/outputs that travel to the GPU box.
The six SURVEY Appendix B positions are included so the file also carries the
surveyor's independently derived values.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib  # noqa: E402
from squava_b200 import positions  # noqa: E402

APPENDIX_B = [(0x1080038, 0x0c00181), (0x0504208, 0x00280a1), (0x0c0090a, 0x02400d4),
              (0x1408051, 0x0a21500), (0x05c2408, 0x1800274), (0x0412825, 0x098c0c0)]


def main():
    o = oracle_lib.load()
    t = o.tables()
    out = {"about": "C-oracle outputs (oracle/tools/make_golden.py); PARITY UNPINNED vs reference output"}
    g = np.random.Generator(np.random.PCG64(2026))
    ab1 = []
    cases = [(x, om, 10) for x, om in APPENDIX_B]
    for marks, depth, count in ((10, 10, 3), (12, 10, 4), (14, 10, 4), (16, 12, 3), (8, 6, 2), (4, 4, 2), (0, 3, 1)):
        for _ in range(count):
            x, om = positions.random_position(g, marks)
            cases.append((x, om, depth))
    for x, om, depth in cases:
        vals, bm, bv, lv = o.ab1_root(x, om, depth, t["scores_ab"])
        ab1.append({"x": x, "o": om, "depth": depth, "values": vals, "best_mask": bm, "best_value": bv, "leaves": lv})
    out["ab1_root"] = ab1
    ab2 = []
    for marks, depth, count in ((10, 9, 2), (12, 11, 3), (14, 12, 3), (16, 15, 2), (6, 5, 2)):
        for _ in range(count):
            x, om = positions.random_position(g, marks)
            vals, bm, bv, lv = o.ab2_root(x, om, depth, t["scores_ab"])
            ab2.append({"x": x, "o": om, "depth": depth, "values": vals, "best_mask": bm, "best_value": bv, "leaves": lv})
    out["ab2_root"] = ab2
    # opening3 / opening2 at reduced depth (the programs' own subtrees)
    op3 = []
    for depth in (6, 7):
        for f, r in o.opening3_subtrees():
            v, lv = o.ab_subtree(3, 1 << f, 1 << r, r, 2, 1, t["scores_opening"][f], depth, t["scores_opening"])
            op3.append({"first": f, "reply": r, "depth": depth, "value": v, "leaves": lv})
    out["opening3"] = op3
    op2 = []
    for depth in (6, 8):
        for f in t["first_moves"]:
            v, lv = o.ab_subtree(4, 1 << f, 0, f, 1, -1, 0, depth, t["scores_opening"])
            op2.append({"first": f, "depth": depth, "value": v, "leaves": lv})
    out["opening2"] = op2
    ex = []
    for marks in (10, 11, 12, 13, 14, 16):
        x, om = positions.random_position(g, marks)
        tm = 1 if marks % 2 == 0 else -1
        p, plies = o.playout_exact(x, om, tm)
        ex.append({"x": x, "o": om, "to_move": tm, "p": p, "mean_plies": plies})
    for x, om in APPENDIX_B:
        p, plies = o.playout_exact(x, om, 1)
        ex.append({"x": x, "o": om, "to_move": 1, "p": p, "mean_plies": plies})
    out["playout_exact"] = ex
    path = os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
