#!/usr/bin/env python3
"""Timing of the alpha-beta batch (configs[3]-style inputs).  Not the headline bench.

usage: bench_ab.py [--marks 8,10,12,14] [--count 64] [--depth 10] [--dialect 1] [--check K]
--check K: compare the first K positions of every group with the CPU oracle (values, tie sets)
and report the oracle's reference-order leaf counts and its single-core rate.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--marks", default="14,12,10,8")
    ap.add_argument("--count", type=int, default=64)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--dialect", type=int, default=1)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--seed", type=int, default=404)
    args = ap.parse_args()
    import squava_b200 as sq
    from squava_b200 import positions
    import oracle_lib
    o = oracle_lib.load()
    sc = o.tables()["scores_ab"]
    sq.init()
    sq.alphabeta_root(positions.midgame_batch(4, marks=(14,), seed=1)[0], args.dialect, 4, sc)   # warm up
    for m in [int(v) for v in args.marks.split(",")]:
        boards, _ = positions.midgame_batch(args.count, marks=(m,), seed=args.seed + m)
        if m % 2:
            continue
        l0 = sq.launch_count()
        t0 = time.perf_counter()
        values, bv, bm, nodes = sq.alphabeta_root(boards, args.dialect, args.depth, sc)
        dt = time.perf_counter() - t0
        rec = {"marks": m, "positions": args.count, "depth": args.depth, "dialect": args.dialect,
               "seconds": dt, "gpu_evals": int(nodes.sum()), "gpu_evals_per_s": float(nodes.sum() / dt),
               "launches": sq.launch_count() - l0, "split": os.environ.get("SQ_AB_SPLIT", "auto")}
        if args.check:
            k = min(args.check, args.count)
            t0 = time.perf_counter()
            leaves = 0
            oevals = 0
            for i in range(k):
                f = o.ab1_root if args.dialect == 1 else o.ab2_root
                ov, obm, obv, lv = f(int(boards[i]["x"]), int(boards[i]["o"]), args.depth, sc)
                assert values[i].tolist() == ov and int(bm[i]) == obm and int(bv[i]) == obv, (m, i)
                leaves += lv
                oevals += o.last_evals()
            cdt = time.perf_counter() - t0
            rec.update({"checked": k, "oracle_leaves_checked": leaves, "oracle_evals_checked": oevals, "oracle_seconds_1core": cdt,
                        "oracle_leaves_per_s_1core": leaves / max(cdt, 1e-9),
                        "gpu_evals_checked": int(nodes[:k].sum())})
        print(json.dumps(rec), flush=True)
    sq.shutdown()


if __name__ == "__main__":
    main()
