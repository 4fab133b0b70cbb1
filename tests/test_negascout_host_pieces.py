"""The host-side pieces the NegaScout split driver (NsSplit, squava_b200/csrc/abi.cu) rests on -- the children of a node
in the mover's order (reorderMoves, negascout.go:411-499) and "does this move end the game" (the driver keeps finished
games as items) -- compiled for the host from the SAME header the library uses and compared with the Python oracle on
random boards.  No GPU needed."""
import os
import random
import shutil
import subprocess

import pytest

import pyoracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_made(bd, c, pl):
    """pl's mark on the empty cell c completes a four or a three of pl's"""
    after = list(bd)
    after[c] = pl
    return any(all(after[k] == pl for k in q) for q in P.IQ[c]) or any(all(after[k] == pl for k in t) for t in P.IT[c])


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host-only compile)")
def test_children_order_and_game_ending_moves(tmp_path):
    exe = str(tmp_path / "ns_host_check")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "ns_host_check.cu")])
    rng = random.Random(12)
    boards = []
    for _ in range(3000):
        marks = rng.randrange(0, 24)
        cells = rng.sample(range(25), marks)
        bd = [P.UNSET] * 25
        for i, c in enumerate(cells):
            bd[c] = 1 if (i % 2 == 0) == (rng.random() < 0.9) else -1        # mostly alternating, some lopsided boards
        boards.append(bd)
    text = "".join("%d %d\n" % (sum(1 << c for c in range(25) if bd[c] == 1), sum(1 << c for c in range(25) if bd[c] == -1))
                   for bd in boards)
    out = subprocess.run([exe], input=text, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    lines = out.stdout.strip().split("\n")
    assert len(lines) == len(boards)
    ending = 0
    for bd, line in zip(boards, lines):
        f = line.split()
        n = int(f[0])
        order_min, order_max = P.ns_reorder(bd)
        assert n == len(order_max) == sum(1 for v in bd if v == P.UNSET)
        assert [int(v) for v in f[1:1 + n]] == order_max
        assert [int(v) for v in f[1 + n:1 + 2 * n]] == order_min
        for item in f[1 + 2 * n:]:
            c, ab = item.split(":")
            want = "%d%d" % (line_made(bd, int(c), 1), line_made(bd, int(c), -1))
            assert ab == want, (bd, c)
            ending += ab != "00"
    assert ending > 2000
