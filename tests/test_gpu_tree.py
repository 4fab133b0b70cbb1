"""sq_mcts_decide: one UCT decision with the tree in HBM (squava_b200/csrc/tree.cu).  Rollouts are random, so the checks
are the ones the host arena tree gets (tests/test_host_uct_vs_reference.py): the books balance, the search finds what
there is to find, and it spends its budget like the host tree does on the same position."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(sq):
    sq.init()
    yield sq
    sq.shutdown()


def test_books_balance(dev):
    for iters, batch, ppl, walkers, levels in ((300_000, 4096, 1, 0, -1), (1_000_000, 65536, 1, 0, -1), (400_000, 1000, 4, 256, 2),
                                               (50_001, 1, 1, 1, 3), (200_000, 65536, 1, 65536, 0), (300_000, 4096, 1, 0, 9)):
        r = dev.mcts_decide(0, 0, -1, iters, batch=batch, playouts_per_leaf=ppl, walkers=walkers, sync_levels=levels, seed=iters)
        assert r["iterations"] == iters and r["root_visits"] == iters and not r["arena_full"]
        assert sorted(r["moves"]) == list(range(25))
        # walks that find no child of the root ready roll out from the root itself: the very first batch (256 walks), and
        # without synchronised levels whoever arrives while the first children are being set up
        assert iters - 4096 * ppl <= r["visits"].sum() <= iters
        assert (r["wins"] >= 0).all() and (r["wins"] <= r["visits"]).all()
        # (without synchronised levels a batch whose walks all run at once follows a few lines only: the tree stays small)
        assert r["nodes"] > (iters // ppl if levels != 0 else 1000) and r["best_move"] in range(25)


def test_finds_the_move_that_wins_at_once_and_the_one_that_must_be_blocked(dev):
    x, o = (1 << 0) | (1 << 1) | (1 << 3), (1 << 10) | (1 << 12) | (1 << 19)
    r = dev.mcts_decide(x, o, -1, 200_000, batch=4096, seed=3)                         # X to move: cell 2 makes four
    k = r["moves"].index(2)
    assert r["best_move"] == 2 and r["visits"][k] > 0.6 * 200_000 and r["wins"][k] > 0.95 * r["visits"][k]
    r = dev.mcts_decide(x | (1 << 24), o, 1, 400_000, batch=4096, seed=4)              # O to move: block it or lose
    assert r["moves"][int(np.argmax(r["visits"]))] == 2
    full = dev.mcts_decide(0x1364d93, 0x1ffffff & ~0x1364d93, 1, 1000, batch=64)        # no empty cell: nothing to grow
    assert full["root_visits"] == 1000 and full["moves"] == [] and full["best_move"] is None


def test_small_arena_stops_growing_but_keeps_counting(dev):
    r = dev.mcts_decide(0, 0, -1, 300_000, batch=4096, max_nodes=20_000, seed=9)
    assert r["arena_full"] and r["root_visits"] == 300_000 and r["nodes"] >= 20_000 - 25
    assert (r["wins"] <= r["visits"]).all() and 300_000 - 4096 <= r["visits"].sum() <= 300_000


def test_spends_its_budget_like_the_host_tree(dev):
    """the same decision by the host arena tree (squava_b200/host/fast_tree.cpp, real GPU rollouts) and by the tree in
    HBM: both put most of their visits on the same handful of first moves"""
    L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "squava_b200", "host", "libsquava_host.so"))
    L.sqh_fast_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                               C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                               C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    x, o = 0, 0                                                                          # the empty board, X to move
    iters = 2_000_000
    moves = (C.c_int * 25)(); visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
    nch, best, value, rootv = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    rc = L.sqh_fast_uct(x, o, -1, iters, 1.0, 1, 8192, 1, 4, None, None, None, None, C.byref(nch), moves, visits, wins,
                        C.byref(best), C.byref(value), C.byref(rootv), None, None)
    assert rc == 0
    host = np.zeros(25)
    for i in range(nch.value):
        host[moves[i]] = visits[i]
    r = dev.mcts_decide(x, o, -1, iters, batch=8192, seed=11)
    gpu = np.zeros(25)
    for m, v in zip(r["moves"], r["visits"]):
        gpu[m] = v
    host /= host.sum(); gpu /= gpu.sum()
    # the board is symmetric: WHICH corner a search ends up liking is a coin toss, how much of the budget the corners, the
    # edges, the inner ring and the centre get is not
    classes = {"corners": [0, 4, 20, 24], "centre": [12], "ring": [6, 7, 8, 11, 13, 16, 17, 18],
               "edges": [1, 2, 3, 5, 9, 10, 14, 15, 19, 21, 22, 23]}
    share = {name: (float(host[cells].sum()), float(gpu[cells].sum())) for name, cells in classes.items()}
    assert max(share, key=lambda n: share[n][0]) == max(share, key=lambda n: share[n][1]), share
    assert all(abs(h - g) < 0.3 for h, g in share.values()), share
