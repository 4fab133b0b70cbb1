"""The HBM tree's walk and update (squava_b200/csrc/tree.cuh: the __host__ __device__ functions the select and update
kernels call) run sequentially on the host against a CPU rollout (tests/native/tree_check.cu): every iteration is
counted once at every level, wins never exceed visits, a move that wins at once takes the visits, a finished root
grows nothing.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(exe, *args):
    out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    f = lines[0].split()
    head = {f[i]: int(f[i + 1]) for i in range(0, len(f), 2)}
    kids = {int(l.split()[1]): (int(l.split()[2]), int(l.split()[3])) for l in lines[1:]}
    return head, kids


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host-only compile)")
def test_tree_walk_and_update_keep_their_books(tmp_path):
    exe = str(tmp_path / "tree_check")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "tree_check.cu")])
    for levels in (0, 1, 4, 9):
        for iters, batch, ppl, seed in ((200_000, 4096, 1, 7), (300_000, 2000, 4, 5), (50_001, 1, 1, 3), (10_000, 65536, 1, 2)):
            head, kids = run(exe, iters, batch, ppl, seed, 0, 0, -1, levels)
            assert head["root_visits"] == iters == head["iterations"] and head["bad"] == 0 and head["overflow"] == 0
            # synchronised levels: the walks of the very first batch (256) that find no child ready roll out from the root
            gap = 0 if levels == 0 or batch == 1 else (min(256, batch) - 25) * ppl
            assert len(kids) == 25 and sum(v for v, _ in kids.values()) == iters - gap == iters - head["root_gap"]
            assert all(0 <= w <= v for v, w in kids.values())
    # the synchronised descent shares a batch out like sequential selections with virtual loss: same favourite first moves
    seq = run(exe, 600_000, 16384, 1, 7, 0, 0, -1, 0)[1]
    syn = run(exe, 600_000, 16384, 1, 7, 0, 0, -1, 4)[1]
    top = lambda kids: set(sorted(kids, key=lambda c: -kids[c][0])[:6])
    assert len(top(seq) & top(syn)) >= 4 and max(v for v, _ in syn.values()) < 0.3 * 600_000
    # X holds 0, 1 and 3 of the top row: cell 2 makes four at once (three of the same row would LOSE: four is checked first)
    x, o = (1 << 0) | (1 << 1) | (1 << 3), (1 << 10) | (1 << 12) | (1 << 19)
    for levels in (0, 4):
        head, kids = run(exe, 100_000, 1024, 1, 3, x, o, -1, levels)
        assert max(kids, key=lambda c: kids[c][0]) == 2 and kids[2][0] > 50_000 and kids[2][1] > 0.95 * kids[2][0]
    # the same position with O to move: O must take cell 2 or lose -- nothing else gets the visits
    for levels in (0, 4):
        head, kids = run(exe, 200_000, 1024, 1, 4, x | (1 << 24), o, 1, levels)
        assert max(kids, key=lambda c: kids[c][0]) == 2
    # a full board: nothing to grow, every walk ends at the root
    head, kids = run(exe, 1000, 64, 1, 5, 0x1ffffff, 0, 1)
    assert head["root_visits"] == 1000 and head["nodes"] == 1 and not kids
