#!/usr/bin/env python3
"""Per-call latency of the host-buffer entry points at MCTS-sized batches."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import squava_b200 as sq
from squava_b200 import positions
sq.init()
out = {}
for n, ppl in ((1, 1), (256, 1), (4096, 1), (65536, 1), (256, 64), (65536, 64)):
    b, tm = positions.midgame_batch(min(n, 4096), seed=1)
    b = np.resize(b, n); tm = np.resize(tm, n)
    res = np.empty((n, 4), np.uint64)
    for _ in range(20):
        sq.playouts(b, tm, ppl, 0, out=res)
    reps = 300 if n <= 4096 else 50
    t0 = time.perf_counter()
    for i in range(reps):
        sq.playouts(b, tm, ppl, i, out=res)
    dt = (time.perf_counter() - t0) / reps
    out["%dx%d" % (n, ppl)] = {"us_per_call": dt * 1e6, "playouts_per_s": n * ppl / dt}
f = np.zeros(1, sq.BOARD)
t0 = time.perf_counter()
for i in range(300):
    sq.classify(f)
out["classify_1"] = {"us_per_call": (time.perf_counter() - t0) / 300 * 1e6}
print(json.dumps(out))
sq.shutdown()
