#!/usr/bin/env python3
"""Exhaustive parity sweep of the terminal-position test: EVERY board with at
most K marks and X-count - O-count in {0, 1} (K=10: 1,177,833,846 boards), GPU through the C ABI
vs the CPU oracle, compared element-wise.  The oracle side (enumeration by rank + table-walk
classification) runs on a thread pool; ctypes releases the GIL.

usage: sweep_classify_exhaustive.py [K=10] [threads=nproc] [Kmin=0]   (marks Kmin..K)"""
import ctypes as C
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import squava_b200 as sq
import oracle_lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
T = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else (os.cpu_count() or 1)
KMIN = int(sys.argv[3]) if len(sys.argv) > 3 else 0
CH = 1 << 21
o = oracle_lib.load()
L = o.L
L.sqo_count_alternating.restype = C.c_int64
L.sqo_count_alternating.argtypes = [C.c_int]
L.sqo_enumerate_alternating.restype = None
L.sqo_enumerate_alternating.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p]
sq.init()
gpu_lock = threading.Lock()


def chunk(job):
    k, start, n = job
    b = np.empty(n, sq.BOARD)
    L.sqo_enumerate_alternating(k, start, n, b.ctypes.data)
    of, owm, owa = o.classify(b)
    with gpu_lock:
        f, wm, wa = sq.classify(b)
    bad = int((f != of).sum() + (wm != owm).sum() + (wa != owa).sum())
    pc = np.bitwise_count(b["x"]).astype(np.int64) - np.bitwise_count(b["o"])
    assert ((pc == 0) | (pc == 1)).all() and not (b["x"] & b["o"]).any()
    return n, bad, int((f != 0).sum()), int((wm != wa).sum()), int((f.astype(np.uint64) * 3 + (wm + 1) * 5 + (wa + 1) * 7).sum())


jobs, per_k = [], {}
for k in range(KMIN, K + 1):
    tot = L.sqo_count_alternating(k)
    per_k[k] = tot
    for s in range(0, tot, CH):
        jobs.append((k, s, min(CH, tot - s)))
t0 = time.time()
boards = mism = term = both = chk = 0
with ThreadPoolExecutor(T) as ex:
    for n, bad, tm, bo, c in ex.map(chunk, jobs):
        boards += n; mism += bad; term += tm; both += bo; chk += c
print(json.dumps({"min_marks": KMIN, "max_marks": K, "boards": boards, "per_marks": per_k, "mismatches": mism, "terminal": term,
                  "scan_orders_disagree_on": both, "checksum": chk, "threads": T,
                  "seconds": round(time.time() - t0, 1)}))
sq.shutdown()
sys.exit(0 if mism == 0 else 1)
