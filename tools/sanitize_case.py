#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the library once,
tiny sizes, with the alpha-beta budget forced down so that the adaptive splitting (expand, brothers waiting and
being set free, the lock-free completion protocol) is exercised too.

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py     (shared-memory hazards: the playout kernel's
                                                                          Fisher-Yates array, the search stacks)
One tool per gpurun call (B200_PROFILING.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import squava_b200 as sq
from squava_b200 import positions

sq.init()
b, tm = positions.midgame_batch(40, marks=(12, 14, 16), seed=2)
print(sq.classify(b)[0][:4])
print(sq.playouts(b, tm, 70, 1)[:2])
print(sq.playouts_packed(b, tm, 3)[:8])
sc = np.array([3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3], np.int8)
os.environ["SQ_AB_BUDGET"] = "16"
for d in (1, 2, 5, 6):
    print(sq.alphabeta_root(b[:6], d, 6, sc)[1])
t = np.zeros(2, sq.SUBTREE)
t[0] = (1 << 12, 1 << 6, 6, 2, 1, 0, 0)
t[1] = (1 << 0, 0, 0, 1, -1, 0, 0)
print(sq.alphabeta_subtrees(t[:1], 3, 5, sc), sq.alphabeta_subtrees(t[1:], 4, 4, sc))
r = sq.negascout_root(b[:4], 4, sc)
print(r["best_value"], r["nodes"])
sq.shutdown()
print("done")
