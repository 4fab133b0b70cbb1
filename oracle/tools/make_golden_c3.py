#!/usr/bin/env python3
"""Write tests/golden/c3_eight_mark_depth10.json: the C oracle's src/alphabeta values (dialect AB-1,
maxDepth 10, default bias table: BASELINE configs[3]) on EIGHT-MARK positions -- the 10^7..3*10^8-leaf
trees that drive the GPU search's adaptive splitting and that are too slow to search live in every
test run (1..45 s each on one host core).

TEST INFRASTRUCTURE.  The positions are the first eight-mark positions of the configs[3] batch
generator (squava_b200.positions.midgame_batch(.., marks=(8, 10, 12, 14), seed=SEED_C2)) plus a
seeded extra batch; the oracle itself is pinned to the Go code's output for this dialect in
tests/test_reference_vectors.py.
"""
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib  # noqa: E402
from squava_b200 import positions  # noqa: E402


def main():
    o = oracle_lib.load()
    sc = o.tables()["scores_ab"]
    boards, _ = positions.midgame_batch(1024, marks=(8, 10, 12, 14))
    eight = [(int(b["x"]), int(b["o"])) for b in boards if bin(int(b["x"]) | int(b["o"])).count("1") == 8][:6]
    extra, _ = positions.midgame_batch(4, marks=(8,), seed=0xC3)
    eight += [(int(b["x"]), int(b["o"])) for b in extra]

    def one(xo):
        vals, bm, bv, lv = o.ab1_root(xo[0], xo[1], 10, sc)      # ctypes releases the GIL: one core per position
        return {"x": xo[0], "o": xo[1], "depth": 10, "values": vals, "best_mask": bm, "best_value": bv, "leaves": lv}

    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        cases = list(ex.map(one, eight))
    out = {"about": "C-oracle AB-1 depth-10 root values on eight-mark positions (oracle/tools/make_golden_c3.py)",
           "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "c3_eight_mark_depth10.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, [c["leaves"] for c in cases])


if __name__ == "__main__":
    main()
