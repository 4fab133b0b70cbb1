#!/usr/bin/env python3
"""One process, several GPUs: sq_init(ids, n) shards every host-buffer call over the devices.
Results must equal the one-device call bit for bit.  usage: multi_device_inprocess.py [n_devices]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import squava_b200 as sq
from squava_b200 import positions

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
boards, tm = positions.midgame_batch(4096, seed=33)
ab_boards, _ = positions.midgame_batch(256, marks=(8, 10, 12, 14), seed=34)
sc = np.array([3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3], np.int8)
res = {}
for devs in ([0], list(range(n_dev))):
    sq.init(devices=devs, seed=77)
    sq.playouts(boards[:64], tm[:64], 100, 0)
    t0 = time.perf_counter()
    p = sq.playouts(boards, tm, 100_000, 5)
    dt = time.perf_counter() - t0
    c = sq.classify(boards)
    t0 = time.perf_counter()
    v = sq.alphabeta_root(ab_boards, 1, 10, sc)
    dta = time.perf_counter() - t0
    v6 = sq.alphabeta_root(ab_boards[:64], 6, 8, sc)
    t0 = time.perf_counter()
    ns = sq.negascout_root(ab_boards, 8, sc)
    dtn = time.perf_counter() - t0
    small = sq.playouts(boards[:3], tm[:3], 1000, 9)         # the pinned small-call path, fewer leaves than devices
    res[len(devs)] = (p, c, v[:3] + v6[:3] + (ns["values"], ns["visit_order"], ns["best_mask"], ns["first_best"], ns["nodes"], small), dt, dta, dtn)
    sq.shutdown()
a, b = res[1], res[n_dev]
ok = bool((a[0] == b[0]).all() and all((x == y).all() for x, y in zip(a[1], b[1])) and all((x == y).all() for x, y in zip(a[2], b[2])))
print(json.dumps({"devices": n_dev, "identical_to_one_device": ok, "playouts_per_s_1": 4096 * 1e5 / a[3],
                  "playouts_per_s_n": 4096 * 1e5 / b[3], "alphabeta_s_1": a[4], "alphabeta_s_n": b[4],
                  "negascout_s_1": a[5], "negascout_s_n": b[5]}))
sys.exit(0 if ok else 1)
