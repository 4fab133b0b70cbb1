#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck): every kernel once, tiny sizes."""
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
sc = np.array([3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3], np.int8)
os.environ["SQ_AB_BUDGET"] = "64"
for d in (1, 2):
    print(sq.alphabeta_root(b[:6], d, 6, sc)[1])
t = np.zeros(2, sq.SUBTREE)
t[0] = (1 << 12, 1 << 6, 6, 2, 1, 0, 0)
t[1] = (1 << 0, 0, 0, 1, -1, 0, 0)
print(sq.alphabeta_subtrees(t[:1], 3, 5, sc), sq.alphabeta_subtrees(t[1:], 4, 4, sc))
sq.shutdown()
print("done")
