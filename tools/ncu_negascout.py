#!/usr/bin/env python3
"""One sq_negascout_root batch (4096 positions, 10 marks, depth 8) for an ncu capture of negascout_fast_kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import squava_b200 as sq
from squava_b200 import positions

sc = [3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3]
sq.init()
boards, _ = positions.midgame_batch(4096, marks=(10,), seed=4010)
r = sq.negascout_root(boards, 8, sc)
print("nodes", int(r["nodes"].sum()), "max", int(r["nodes"].max()))
sq.shutdown()
