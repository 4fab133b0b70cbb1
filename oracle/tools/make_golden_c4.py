#!/usr/bin/env python3
"""Write tests/golden/c4_opening3_depth10.json: the C oracle's opening3 values (dialect AB-3, -d 10, the
program's default: BASELINE configs[4]) for all 85 symmetry-unique (first move, reply) subtrees, with the
oracle's sequential evaluator-call and leaf counts (the "reference-order" work the GPU rate is quoted in).

TEST INFRASTRUCTURE.  Minutes of CPU on all cores (about 5*10^10 evaluator calls), which is why the vectors
are committed instead of being recomputed in the tests; the oracle itself is pinned to the Go code's output
for this dialect at depths 4..6 in tests/test_reference_vectors.py.
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

DEPTH = int(os.environ.get("C4_DEPTH", "10"))


def one(fr):
    import oracle_lib
    o = oracle_lib.load()
    sc = o.tables()["scores_opening"]
    f, r = fr
    t0 = time.time()
    v, leaves = o.ab_subtree(3, 1 << f, 1 << r, r, 2, 1, sc[f], DEPTH, sc)
    return {"first": f, "reply": r, "depth": DEPTH, "value": v, "leaves": leaves, "evals": o.last_evals(), "seconds": round(time.time() - t0, 2)}


def main():
    import oracle_lib
    pairs = oracle_lib.load().opening3_subtrees()
    assert len(pairs) == 85
    t0 = time.time()
    with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
        cases = pool.map(one, pairs, chunksize=1)
    out = {"about": "C-oracle opening3 (AB-3) values at depth %d for the 85 subtrees (oracle/tools/make_golden_c4.py)" % DEPTH,
           "wall_seconds": round(time.time() - t0, 1), "procs": len(os.sched_getaffinity(0)),
           "total_evals": sum(c["evals"] for c in cases), "total_leaves": sum(c["leaves"] for c in cases), "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "c4_opening3_depth%d.json" % DEPTH)
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, out["wall_seconds"], out["total_evals"])


if __name__ == "__main__":
    main()
