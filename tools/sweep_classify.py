#!/usr/bin/env python3
"""One-off large parity sweep of the terminal-position test: GPU (through the C ABI) vs the CPU
oracle on N random boards, half uniform-ternary (mostly unreachable, both colours own lines) and
half alternating-count boards with 6..25 marks; compared element-wise in chunks.

usage: sweep_classify.py [N=200000000]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import squava_b200 as sq
import oracle_lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
CH = 4_000_000
o = oracle_lib.load()
sq.init()
rng = np.random.Generator(np.random.PCG64(2026))
w = (1 << np.arange(25, dtype=np.uint32))
done = mism = both = term = 0
t0 = time.time()
while done < N:
    n = min(CH, N - done)
    b = np.empty(n, sq.BOARD)
    if (done // CH) % 2 == 0:
        trits = rng.integers(0, 3, size=(n, 25), dtype=np.uint8)
        b["x"] = ((trits == 1) * w).sum(axis=1, dtype=np.uint32)
        b["o"] = ((trits == 2) * w).sum(axis=1, dtype=np.uint32)
    else:
        k = rng.integers(6, 26, size=n)
        rank = np.argsort(np.argsort(rng.random((n, 25), dtype=np.float32), axis=1), axis=1)
        nx = (k + 1) // 2
        b["x"] = ((rank < nx[:, None]) * w).sum(axis=1, dtype=np.uint32)
        b["o"] = (((rank >= nx[:, None]) & (rank < k[:, None])) * w).sum(axis=1, dtype=np.uint32)
    f, wm, wa = sq.classify(b)
    of, owm, owa = o.classify(b)
    mism += int((f != of).sum() + (wm != owm).sum() + (wa != owa).sum())
    both += int((wm != wa).sum())
    term += int((f != 0).sum())
    done += n
print(json.dumps({"boards": done, "mismatches": mism, "terminal": term, "scan_orders_disagree_on": both,
                  "seconds": time.time() - t0}))
sq.shutdown()
sys.exit(0 if mism == 0 else 1)
