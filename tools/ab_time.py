#!/usr/bin/env python3
"""In-process timings of the alpha-beta batch on the BASELINE-sized inputs (configs[3], opening3 -d 10/12, opening2 -d 12;
AB_DEEP=1 adds opening2 -d 13/14), with a checksum of the values so that runs under different scheduling knobs
(SQ_AB_PASSES, SQ_AB_BUDGET, SQ_AB_MIN_BUDGET, SQ_AB_NO_YBW ...) can be compared at a glance."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import squava_b200 as sq  # noqa: E402
from squava_b200 import positions  # noqa: E402

sq.init()
boards, _ = positions.midgame_batch(1024, marks=(8, 10, 12, 14), seed=0xC4)
pairs, st3 = bench.opening3_subtrees(sq)
st2 = np.zeros(6, sq.SUBTREE)
for i, c in enumerate((0, 1, 2, 6, 7, 12)):
    st2[i] = (1 << c, 0, c, 1, -1, 0, 0)
res = {}


def t(name, f, reps=2):
    best = None
    for r in range(reps):
        t0 = time.perf_counter()
        out = f()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res[name] = round(best, 4)
    return out


v = t("c3", lambda: sq.alphabeta_root(boards, 1, 10, bench.SCORES_AB))
res["c3_evals"] = int(v[3].sum())
res["c3_sum"] = int(v[1].astype(np.int64).sum())
v = t("o3_d10", lambda: sq.alphabeta_subtrees(st3, 3, 10, bench.SCORES_OPENING))
res["o3_d10_sum"] = int(v[0].sum())
v = t("o3_d12", lambda: sq.alphabeta_subtrees(st3, 3, 12, bench.SCORES_OPENING), 1)
res["o3_d12_sum"] = int(v[0].sum())
v = t("o2_d12", lambda: sq.alphabeta_subtrees(st2, 4, 12, bench.SCORES_OPENING), 1)
res["o2_d12_sum"] = int(v[0].sum())
res["o2_d12_evals"] = int(v[1].sum())
if os.environ.get("AB_DEEP"):
    for d in (13, 14):
        v = t("o2_d%d" % d, lambda: sq.alphabeta_subtrees(st2, 4, d, bench.SCORES_OPENING), 1)
        res["o2_d%d_values" % d] = v[0].tolist()
print(json.dumps(res))
sq.shutdown()
