#!/usr/bin/env python3
"""Engine-vs-engine batch: the same N games played one after the other (one position
per library call) and in lockstep (one batched call per ply).
usage: bench_tournament.py [N=32] [first=N] [second=A]"""
import json
import os
import subprocess
import sys
import time

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "squava_b200", "host")
n = sys.argv[1] if len(sys.argv) > 1 else "32"
a = sys.argv[2] if len(sys.argv) > 2 else "N"
b = sys.argv[3] if len(sys.argv) > 3 else "A"
for mode in ([], ["-lockstep"]):
    t = time.time()
    out = subprocess.run([os.path.join(HOST, "playoff5"), "-n", n, "-1", a, "-2", b, "-seed", "1"] + mode,
                         capture_output=True, text=True, check=True).stdout
    dt = time.time() - t
    w = [int(l.split(" ")[6]) for l in out.splitlines()]
    print(json.dumps({"games": int(n), "first": a, "second": b, "mode": "lockstep" if mode else "sequential",
                      "seconds": round(dt, 3), "x_wins": w.count(1), "o_wins": w.count(-1), "cat": w.count(0)}), flush=True)
