#!/usr/bin/env python3
"""NegaScout (src/negascout) on the GPU vs the CPU oracle: staticValue calls per second for one
position (latency case, what playoff5 -1 N does per move) and for batches (one warp per position).

usage: bench_negascout.py [--cpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import squava_b200 as sq
from squava_b200 import positions
import oracle_lib

o = oracle_lib.load()
sc = o.tables()["scores_ab"]
sq.init()
out = {}
sq.negascout_root(positions.midgame_batch(4, marks=(14,), seed=1)[0], 4, sc)      # warm-up
for name, marks, depth, count in [("1 position, 4 marks, depth 8", 4, 8, 1), ("1 position, 6 marks, depth 8", 6, 8, 1),
                                  ("1 position, 2 marks, depth 6", 2, 6, 1), ("64 positions, 6 marks, depth 8", 6, 8, 64),
                                  ("1024 positions, 8 marks, depth 8", 8, 8, 1024),
                                  ("8192 positions, 10 marks, depth 8", 10, 8, 8192)]:
    boards, _ = positions.midgame_batch(count, marks=(marks,), seed=4000 + marks)
    t = time.time(); r = sq.negascout_root(boards, depth, sc); gpu_s = time.time() - t
    nodes = int(r["nodes"].sum())
    rec = {"nodes": nodes, "gpu_s": round(gpu_s, 4), "gpu_nodes_per_s": round(nodes / gpu_s)}
    k = min(count, 8)
    t = time.time(); cpu_nodes = 0
    for i in range(k):
        res = o.negascout_root(int(boards[i]["x"]), int(boards[i]["o"]), depth, sc)
        cpu_nodes += res[5]
        assert res[0] == r["values"][i].tolist() and res[5] == int(r["nodes"][i])
    cpu_s = time.time() - t
    rec.update({"cpu_positions_checked": k, "cpu_1core_nodes_per_s": round(cpu_nodes / max(cpu_s, 1e-9))})
    out[name] = rec
    print(name, rec, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "negascout_bench.json"), "w"), indent=1)
sq.shutdown()
