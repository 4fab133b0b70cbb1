"""The CHECKED build of the library (-DSQ_CHECKED=1: the kernels' own bounds checks on node indices, queue slots,
search-stack frames, Fisher-Yates slots and NegaScout call items) on a small mixed workload and on BASELINE-shaped batches: no check may fire.
This is what stands in for compute-sanitizer, which is closed on the development pool (profiles/r02_sanitizer.md)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "squava_b200", "libsquava_b200_checked.so")

SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import squava_b200 as sq
from squava_b200 import positions
assert sq.lib().sq_abi_checked() == 1
sq.init()
sc = np.array([3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3], np.int8)
b, tm = positions.midgame_batch(512, seed=2)
full = np.array([(0x1364d93, 0x1ffffff & ~0x1364d93), (0, 0)], dtype=sq.BOARD)
b = np.concatenate([b, full]); tm = np.concatenate([tm, np.array([-1, 1], np.int8)])
sq.classify(b)
out = sq.playouts(b, tm, 20000, 1)
assert (out[:, :3].sum(axis=1) == 20000).all()
sq.playouts_packed(b, tm, 3)
ab, _ = positions.midgame_batch(48, marks=(8, 10, 12, 14, 20, 22), seed=4)
for env in ({}, {"SQ_AB_BUDGET": "16"}, {"SQ_AB_QUEUE": "1", "SQ_AB_MIN_BUDGET": "8"}, {"SQ_AB_QUEUE": "1", "SQ_AB_NO_YBW": "1", "SQ_AB_POOL_NODES": "6000"}):
    os.environ.update(env)
    for d in (1, 2, 5, 6):
        sq.alphabeta_root(ab, d, 8 if env else 10, sc)
    t = np.zeros(2, sq.SUBTREE)
    t[0] = (1 << 12, 1 << 6, 6, 2, 1, 0, 0); t[1] = (1 << 0, 0, 0, 1, -1, 0, 0)
    sq.alphabeta_subtrees(t[:1], 3, 7, sc); sq.alphabeta_subtrees(t[1:], 4, 6, sc)
    for k in env:
        os.environ.pop(k)
sq.negascout_root(ab[:8], 6, sc)
for env in ({"SQ_NS_SPLIT": "0"}, {"SQ_NS_DIRECT": "0", "SQ_NS_BUDGET": "500", "SQ_NS_ITEM_BUDGET": "100"}, {"SQ_NS_DIRECT": "1000", "SQ_NS_ITEM_BUDGET": "60", "SQ_NS_SPLIT_PLY": "1"}):
    os.environ.update(env)
    sq.negascout_root(ab, 7, sc)
    for k in env:
        os.environ.pop(k)
print("CHECKS", sq.check_failures())
sq.shutdown()
"""


def test_no_in_kernel_check_fires():
    if not os.path.exists(CHECKED):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "squava_b200", "csrc"), "checked"])
    env = dict(os.environ, SQUAVA_B200_LIB=CHECKED)
    p = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("CHECKS")][0]
    assert line == "CHECKS (0, 0, 0, 0)", line
