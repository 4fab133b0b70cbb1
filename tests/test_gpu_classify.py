"""Terminal-position test on the GPU vs the CPU oracle: bit-exact."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(sq):
    sq.init()
    yield sq
    sq.shutdown()


def check(dev, oracle, boards):
    f, wm, wa = dev.classify(boards)
    of, owm, owa = oracle.classify(boards)
    assert (f == of).all() and (wm == owm).all() and (wa == owa).all()
    return f, wm, wa


def alternating_boards(max_marks):
    """every board with X-count - O-count in {0, 1} and at most max_marks marks"""
    out = []
    for k in range(max_marks + 1):
        nx = (k + 1) // 2
        for cells in itertools.combinations(range(25), k):
            for xs in itertools.combinations(cells, nx):
                x = sum(1 << c for c in xs)
                o = sum(1 << c for c in cells) & ~x
                out.append((x, o))
    return np.array(out, dtype=np.uint32)


def test_exhaustive_small(dev, oracle):
    """every reachable-shaped board (X-count - O-count in {0,1}) with at most 6 marks"""
    b = alternating_boards(6).view(dev.BOARD).reshape(-1)
    assert b.shape[0] == 4_156_726
    f, wm, wa = check(dev, oracle, b)
    assert (f != 0).sum() > 0              # with 3+3 marks both colours can own a 3-line: the slow path runs too


def test_random_alternating_dense_boards(dev, oracle):
    """8..24 marks, X-count - O-count in {0,1}: the shapes play produces (lines allowed)"""
    rng = np.random.Generator(np.random.PCG64(10))
    n = 3_000_000
    k = rng.integers(8, 25, size=n)
    order = np.argsort(rng.random((n, 25)), axis=1)
    rank = np.empty_like(order); np.put_along_axis(rank, order, np.arange(25)[None, :].repeat(n, 0), axis=1)
    w = (1 << np.arange(25, dtype=np.uint32))
    nx = (k + 1) // 2
    b = np.empty(n, dev.BOARD)
    b["x"] = ((rank < nx[:, None]) * w).sum(axis=1, dtype=np.uint32)
    b["o"] = (((rank >= nx[:, None]) & (rank < k[:, None])) * w).sum(axis=1, dtype=np.uint32)
    check(dev, oracle, b)


def test_all_full_boards(dev, oracle):
    xs = np.fromiter((sum(1 << c for c in comb) for comb in itertools.combinations(range(25), 13)),
                     dtype=np.uint32, count=5200300)
    b = np.empty(xs.shape[0], dev.BOARD)
    b["x"] = xs
    b["o"] = 0x1ffffff & ~xs
    f, wm, wa = check(dev, oracle, b)
    draws = xs[(f == 16)]
    assert sorted(int(v) for v in draws) == [0x1364d93, 0x1552ab5, 0x15aa955, 0x19364d9]
    assert (wm[f == 16] == 0).all()


def test_random_ternary_boards_including_unreachable(dev, oracle):
    rng = np.random.Generator(np.random.PCG64(9))
    n = 4_000_000
    trits = rng.integers(0, 3, size=(n, 25), dtype=np.uint8)
    w = (1 << np.arange(25, dtype=np.uint32))
    b = np.empty(n, dev.BOARD)
    b["x"] = ((trits == 1) * w).sum(axis=1, dtype=np.uint32)
    b["o"] = ((trits == 2) * w).sum(axis=1, dtype=np.uint32)
    f, wm, wa = check(dev, oracle, b)
    assert (wm != wa).sum() > 0          # scan-order slow path exercised


def test_edge_cases(dev, oracle):
    f, wm, wa = dev.classify(np.zeros(0, dev.BOARD))
    assert f.shape == (0,)
    one = np.zeros(1, dev.BOARD)
    f, wm, wa = dev.classify(one)
    assert f[0] == 0 and wm[0] == 0 and wa[0] == 0
    four = np.array([(0b1111, 0)], dtype=dev.BOARD)
    f, wm, wa = dev.classify(four)
    assert f[0] == 5 and wm[0] == 1 and wa[0] == 1     # four beats the threes it contains
    bad = np.array([(3, 1)], dtype=dev.BOARD)          # overlapping masks are still classified
    dev.classify(bad)
