#!/usr/bin/env python3
"""configs[3]: "squavathr-style root-move alpha-beta at depth 10 over 1024 positions, bit-exact
values".  1024 reachable positions with 8/10/12/14 marks (MAX to move), dialect AB-1
(src/alphabeta), maxDepth 10, default bias table.  GPU batch vs the CPU oracle on all host
cores; EVERY value row and tie mask is compared.  Throughput is reported in reference-order
evaluator calls (the oracle's sequential count) per second, SURVEY.md 8(d).

usage: bench_c4.py [--positions 1024] [--depth 10] [--dialect 1] [--procs N] [--skip-oracle]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def oracle_chunk(args):
    import oracle_lib
    pairs, depth, dialect = args
    o = oracle_lib.load()
    sc = o.tables()["scores_ab"]
    out = []
    for x, om in pairs:
        f = o.ab1_root if dialect == 1 else o.ab2_root
        vals, bm, bv, lv = f(x, om, depth, sc)
        out.append((vals, bm, bv, lv, o.last_evals()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=1024)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--dialect", type=int, default=1)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--skip-oracle", action="store_true")
    args = ap.parse_args()
    import squava_b200 as sq
    from squava_b200 import positions
    import oracle_lib
    sc = oracle_lib.load().tables()["scores_ab"]
    boards, _ = positions.midgame_batch(args.positions, marks=(8, 10, 12, 14), seed=0xC4)
    sq.init()
    sq.alphabeta_root(boards[:8], args.dialect, 4, sc)      # warm-up
    l0 = sq.launch_count()
    t0 = time.perf_counter()
    values, bv, bm, nodes = sq.alphabeta_root(boards, args.dialect, args.depth, sc)
    gpu_s = time.perf_counter() - t0
    rec = {"config": "configs[3]", "positions": args.positions, "depth": args.depth, "dialect": args.dialect,
           "gpu_seconds": gpu_s, "gpu_evaluator_calls": int(nodes.sum()), "gpu_launches": sq.launch_count() - l0,
           "marks_histogram": {str(m): int(sum(1 for b in boards if bin(int(b["x"]) | int(b["o"])).count("1") == m))
                               for m in (8, 10, 12, 14)}}
    sq.shutdown()
    if not args.skip_oracle:
        pairs = [(int(b["x"]), int(b["o"])) for b in boards]
        order = sorted(range(len(pairs)), key=lambda i: bin(pairs[i][0] | pairs[i][1]).count("1"))   # big first
        chunks = [[order[i]] for i in range(len(order))]
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(args.procs) as pool:
            res = pool.map(oracle_chunk, [([pairs[i] for i in c], args.depth, args.dialect) for c in chunks], chunksize=1)
        cpu_s = time.perf_counter() - t0
        mism, evals, leaves = 0, 0, 0
        for c, r in zip(chunks, res):
            for i, (vals, obm, obv, lv, ev) in zip(c, r):
                evals += ev
                leaves += lv
                if values[i].tolist() != vals or int(bm[i]) != obm or int(bv[i]) != obv:
                    mism += 1
        rec.update({"oracle_seconds": cpu_s, "oracle_procs": args.procs, "oracle_evaluator_calls": evals,
                    "oracle_leaves": leaves, "mismatching_positions": mism,
                    "reference_order_evals_per_s_gpu": evals / gpu_s, "reference_order_evals_per_s_cpu": evals / cpu_s,
                    "speedup_vs_cpu_all_cores": cpu_s / gpu_s})
    print(json.dumps(rec), flush=True)
    return 0 if rec.get("mismatching_positions", 0) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
