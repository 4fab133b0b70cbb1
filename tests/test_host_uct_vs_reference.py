"""Host-side UCT (select / expand / backprop / UCB1) against the reference's own driver.

tests/golden/ref_go_uct_vectors.json holds trees grown by src/mcts/mcts.go's and squavam.go's UCT
(translated in memory by oracle/tools/go2py.py, executed by make_ref_golden_uct.py) under a
deterministic rand.Intn.  Fed the same stream -- and a rollout that draws from it exactly like the
reference's loop (mcts.go:173-181) -- the C++ host driver (squava_b200/host, through
libsquava_host.so's callback entry point) and the Python mirror (squava_b200/mcts.py) must grow the
SAME tree: every root child's move, visits and wins, the returned move and int(1000*UCB1).
Sequential mode (batch 1, one playout per leaf).  No GPU needed: the rollout is the test's."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import playout_emulator as E

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
V = json.load(open(os.path.join(HERE, "golden", "ref_go_uct_vectors.json")))
MASK = (1 << 64) - 1


def splitmix(k):
    z = (k * 0x9E3779B97F4A7C15) & MASK
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & MASK
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


class Stream:
    def __init__(self):
        self.k = 0

    def intn(self, n):
        self.k += 1
        return splitmix(self.k) % n

    integers = intn            # numpy-Generator-like face for squava_b200.mcts


def terminal(x, o):
    return E.owns(x, E.TRIPLETS) or E.owns(o, E.TRIPLETS) or (x | o) == 0x1ffffff


def rollout(stream, x, o, to_move):
    """mcts.go:173-181: m := moves[rand.Intn(len(moves))] until terminal.  -> [X wins, O wins, draws, plies]"""
    plies, mover = 0, to_move
    while not terminal(x, o):
        moves = [c for c in range(25) if not ((x | o) >> c) & 1]
        m = moves[stream.intn(len(moves))]
        if mover > 0:
            x |= 1 << m
        else:
            o |= 1 << m
        mover, plies = -mover, plies + 1
    w = 1 if E.owns(x, E.QUADS) else -1 if E.owns(o, E.QUADS) else -1 if E.owns(x, E.TRIPLETS) else 1 if E.owns(o, E.TRIPLETS) else 0
    return [int(w == 1), int(w == -1), int(w == 0), plies]


@pytest.fixture(scope="module")
def host_lib():
    so = os.path.join(ROOT, "squava_b200", "host", "libsquava_host.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.dirname(so), "-s"])
    return C.CDLL(so)


INTN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p)
TERM = C.CFUNCTYPE(C.c_int, C.c_uint32, C.c_uint32, C.c_void_p)
PLAY = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p)


def run_cpp(lib, case):
    st = Stream()

    def playouts(leaves, to_move, n, per_leaf, stream_id, out, ud):
        b = np.ctypeslib.as_array(C.cast(leaves, C.POINTER(C.c_uint32)), shape=(n, 2))
        t = np.ctypeslib.as_array(C.cast(to_move, C.POINTER(C.c_int8)), shape=(n,))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint64)), shape=(n, 4))
        for i in range(n):
            tot = [0, 0, 0, 0]
            for _ in range(per_leaf):
                r = rollout(st, int(b[i, 0]), int(b[i, 1]), int(t[i]))
                tot = [a + c for a, c in zip(tot, r)]
            o[i] = tot
    cb = (INTN(lambda n, ud: st.intn(n)), TERM(lambda x, o, ud: int(terminal(x, o))), PLAY(playouts))
    moves = (C.c_int * 25)(); visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
    nch, best, value, leaves = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.sqh_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, INTN, TERM,
                            PLAY, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                            C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = lib.sqh_uct(case["x"], case["o"], case["pjm"], case["itermax"], case["uctk"], int(case["which"] == "mcts"), 1, 1,
                     cb[0], cb[1], cb[2], None, C.byref(nch), moves, visits, wins, C.byref(best), C.byref(value), C.byref(leaves))
    assert rc == 0
    kids = [[moves[i], visits[i], wins[i]] for i in range(nch.value)]
    return kids, best.value, value.value, st.k


def test_cpp_host_uct_grows_the_reference_tree(host_lib):
    for case in V["cases"]:
        kids, best, value, draws = run_cpp(host_lib, case)
        assert kids == case["children"], (case["which"], hex(case["x"]), case["itermax"])
        assert (best, value, draws) == (case["best"], case["value"], case["draws"])


def test_python_mirror_grows_the_reference_tree():
    from squava_b200 import mcts
    for case in V["cases"]:
        st = Stream()

        def roll(boards, to_move, ppl, stream_id):
            return [rollout(st, x, o, tm) for (x, o), tm in zip(boards, to_move)]
        root = mcts.uct(case["x"], case["o"], case["pjm"], case["itermax"], case["uctk"], roll, st, batch=1, ppl=1,
                        factor2=case["which"] == "mcts")
        kids = [[c.move, c.visits, c.wins] for c in root.children]
        assert kids == case["children"], (case["which"], hex(case["x"]), case["itermax"])
        best = root.best_move(case["uctk"], case["which"] == "mcts")
        assert best.move == case["best"] and st.k == case["draws"]
        assert int(1000. * best.ucb1(case["uctk"], case["which"] == "mcts")) == case["value"]


def run_fast(lib, case):
    """the arena tree of fast_tree.cpp (one thread, batch 1) under the same stream"""
    st = Stream()

    def playouts(leaves, to_move, n, per_leaf, stream_id, out, ud):
        b = np.ctypeslib.as_array(C.cast(leaves, C.POINTER(C.c_uint32)), shape=(n, 2))
        t = np.ctypeslib.as_array(C.cast(to_move, C.POINTER(C.c_int8)), shape=(n,))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint64)), shape=(n, 4))
        for i in range(n):
            o[i] = rollout(st, int(b[i, 0]), int(b[i, 1]), int(t[i]))
    cb = (INTN(lambda n, ud: st.intn(n)), TERM(lambda x, o, ud: int(terminal(x, o))), PLAY(playouts))
    moves = (C.c_int * 25)(); visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
    nch, best, value, rootv = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    lib.sqh_fast_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                 INTN, TERM, PLAY, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    rc = lib.sqh_fast_uct(case["x"], case["o"], case["pjm"], case["itermax"], case["uctk"], int(case["which"] == "mcts"), 1, 1, 1,
                          cb[0], cb[1], cb[2], None, C.byref(nch), moves, visits, wins, C.byref(best), C.byref(value),
                          C.byref(rootv), None, None)
    assert rc == 0
    kids = [[moves[i], visits[i], wins[i]] for i in range(nch.value)]
    return kids, best.value, value.value, st.k, rootv.value


def test_fast_arena_tree_grows_the_reference_tree(host_lib):
    """contiguous child blocks, 32-bit atomic counts, claim-by-fetch_and: still the reference's tree"""
    for case in V["cases"]:
        kids, best, value, draws, rootv = run_fast(host_lib, case)
        assert kids == case["children"], (case["which"], hex(case["x"]), case["itermax"])
        assert (best, value, draws) == (case["best"], case["value"], case["draws"])
        assert rootv == case["root_visits"]


def test_fast_arena_tree_many_threads_keeps_its_books(host_lib):
    """throughput mode (4 threads, groups of interleaved walks, single-precision UCB1, deferred root
    counts) with a synthetic rollout: every iteration is counted exactly once at every level"""
    rng = np.random.Generator(np.random.PCG64(3))
    tr = np.array([sum(1 << c for c in t) if not isinstance(t, int) else t for t in E.TRIPLETS], dtype=np.uint32)

    def playouts(leaves, to_move, n, per_leaf, stream_id, out, ud):
        b = np.ctypeslib.as_array(C.cast(leaves, C.POINTER(C.c_uint32)), shape=(n, 2))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint64)), shape=(n, 4))
        x, oo = b[:, 0][:, None], b[:, 1][:, None]
        term = ((x & tr) == tr).any(axis=1) | ((oo & tr) == tr).any(axis=1) | ((b[:, 0] | b[:, 1]) == 0x1ffffff)
        r = rng.integers(0, per_leaf + 1, size=n).astype(np.uint64)
        o[:, 0] = r; o[:, 1] = np.uint64(per_leaf) - r; o[:, 2] = 0; o[:, 3] = np.where(term, 0, 3 * per_leaf)
    cb = (TERM(lambda x, o, ud: int(terminal(x, o))), PLAY(playouts))
    lib = host_lib
    lib.sqh_fast_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                 INTN, TERM, PLAY, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    # the third case is squavam -fast's own default budget on a many-core host: every thread claims a private
    # 65536-slot piece of the arena, which the arena must have room for however small the budget is
    for iters, batch, ppl, threads in ((300_000, 4096, 1, 4), (400_000, 1000, 4, 4), (10_000, 256, 1, 16), (10_000, 256, 1, 8)):
        moves = (C.c_int * 25)(); visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
        nch, best, value, rootv, nodes = C.c_int(), C.c_int(), C.c_int(), C.c_double(), C.c_ulonglong()
        rc = lib.sqh_fast_uct(0, 0, 1, iters, 1.0, 0, batch, ppl, threads, INTN(0), cb[0], cb[1], None, C.byref(nch), moves, visits, wins,
                              C.byref(best), C.byref(value), C.byref(rootv), None, C.byref(nodes))
        assert rc == 0 and nch.value == 25 and sorted(moves[i] for i in range(25)) == list(range(25))
        assert rootv.value == iters and sum(visits[i] for i in range(25)) == iters
        assert all(0 <= wins[i] <= visits[i] for i in range(25))
        assert best.value in range(25) and nodes.value > iters // ppl


def test_fast_arena_tree_pipelined_batches_keep_their_books(host_lib, monkeypatch):
    """double-buffered batches: the rollouts of batch k run on a second thread while the host threads select batch k+1
    (they see batch k's visits but not yet its wins); every iteration must still be counted exactly once everywhere"""
    monkeypatch.setenv("SQUAVA_FAST_PIPELINE", "1")
    test_fast_arena_tree_many_threads_keeps_its_books(host_lib)
    monkeypatch.delenv("SQUAVA_FAST_PIPELINE")
    monkeypatch.setenv("SQUAVA_FAST_NO_PIPELINE", "1")
    test_fast_arena_tree_many_threads_keeps_its_books(host_lib)


def test_fast_arena_tree_throughput_mode_searches_like_the_exact_mode(host_lib):
    """same synthetic game (the chance that X wins a rollout depends on who holds the inner ring), same budget:
    the 4-thread / batched / single-precision / sharded search must spend its visits on the same first
    moves as the sequential double-precision one"""
    good = sum(1 << c for c in (6, 7, 8, 11, 13, 16, 17, 18))
    tr = np.array([sum(1 << c for c in t) if not isinstance(t, int) else t for t in E.TRIPLETS], dtype=np.uint32)
    rng = np.random.Generator(np.random.PCG64(12))

    def playouts(leaves, to_move, n, per_leaf, stream_id, out, ud):
        b = np.ctypeslib.as_array(C.cast(leaves, C.POINTER(C.c_uint32)), shape=(n, 2))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint64)), shape=(n, 4))
        x, oo = b[:, 0], b[:, 1]
        term = ((x[:, None] & tr) == tr).any(axis=1) | ((oo[:, None] & tr) == tr).any(axis=1) | ((x | oo) == 0x1ffffff)
        edge = np.bitwise_count(x & np.uint32(good)).astype(np.float64) - np.bitwise_count(oo & np.uint32(good))
        p = 1.0 / (1.0 + np.exp(-0.9 * edge))
        xw = rng.binomial(per_leaf, p).astype(np.uint64)
        o[:, 0] = xw; o[:, 1] = np.uint64(per_leaf) - xw; o[:, 2] = 0; o[:, 3] = np.where(term, 0, 3 * per_leaf)
    lib = host_lib
    lib.sqh_fast_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                 INTN, TERM, PLAY, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    term_cb, play_cb = TERM(lambda x, o, ud: int(terminal(x, o))), PLAY(playouts)
    pick = np.random.Generator(np.random.PCG64(13))
    share = {}
    for mode, (batch, threads, intn) in {"exact": (1, 1, INTN(lambda n, ud: int(pick.integers(n)))), "fast": (2048, 4, INTN(0))}.items():
        moves = (C.c_int * 25)(); visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
        nch, best, value, rootv = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        # MINIMIZER just moved: X (the side the inner ring favours) is to move at the root
        rc = lib.sqh_fast_uct(0, 0, -1, 60_000, 0.7, 1, batch, 1, threads, intn, term_cb, play_cb, None, C.byref(nch), moves, visits,
                              wins, C.byref(best), C.byref(value), C.byref(rootv), None, None)
        assert rc == 0 and nch.value == 25
        v = np.zeros(25)
        for i in range(25):
            v[moves[i]] = visits[i]
        share[mode] = v / v.sum()
        assert good >> best.value & 1, (mode, best.value)             # both settle on an inner-ring cell
    a, b = share["exact"], share["fast"]
    inner = [c for c in range(25) if good >> c & 1]
    # the eight inner-ring cells are worth the same, so WHICH of them gets the most visits is a coin toss (virtual
    # loss spreads a batch over them); what must agree is how much of the budget the ring gets as a whole
    assert a[inner].sum() > 0.6 and b[inner].sum() > 0.6 and abs(a[inner].sum() - b[inner].sum()) < 0.2
    outer = [c for c in range(25) if not good >> c & 1]
    assert a[outer].max() < 0.05 and b[outer].max() < 0.05


def test_root_parallel_searchers_merge_their_root_statistics(host_lib):
    """root-parallel MCTS (N searchers, one decision: squavam2.go:93-112 without the shared tree): every iteration
    is counted once over all searchers, the merged statistics equal the sum of what the searchers report, and the
    merged decision lands where a single tree with the whole budget lands (synthetic game: the inner ring wins)"""
    good = sum(1 << c for c in (6, 7, 8, 11, 13, 16, 17, 18))
    tr = np.array([sum(1 << c for c in t) if not isinstance(t, int) else t for t in E.TRIPLETS], dtype=np.uint32)
    rng = np.random.Generator(np.random.PCG64(21))
    streams = set()

    def playouts(leaves, to_move, n, per_leaf, stream_id, out, ud):
        streams.add(stream_id >> 48)                     # searcher i draws from stream ids (i + 1) << 48 ...
        b = np.ctypeslib.as_array(C.cast(leaves, C.POINTER(C.c_uint32)), shape=(n, 2))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint64)), shape=(n, 4))
        x, oo = b[:, 0], b[:, 1]
        term = ((x[:, None] & tr) == tr).any(axis=1) | ((oo[:, None] & tr) == tr).any(axis=1) | ((x | oo) == 0x1ffffff)
        edge = np.bitwise_count(x & np.uint32(good)).astype(np.float64) - np.bitwise_count(oo & np.uint32(good))
        xw = rng.binomial(per_leaf, 1.0 / (1.0 + np.exp(-0.9 * edge))).astype(np.uint64)
        o[:, 0] = xw; o[:, 1] = np.uint64(per_leaf) - xw; o[:, 2] = 0; o[:, 3] = np.where(term, 0, 3 * per_leaf)
    lib = host_lib
    lib.sqh_root_parallel_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_ulonglong, C.c_ulonglong, TERM, PLAY, C.c_void_p, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]
    term_cb, play_cb = TERM(lambda x, o, ud: int(terminal(x, o))), PLAY(playouts)
    share = {}
    for n_searchers in (1, 4):
        streams.clear()
        visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
        best, value, rootv, secs = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        rc = lib.sqh_root_parallel_uct(0, 0, -1, 120_000, 0.7, 1, 1024, 2, 4, n_searchers, -1, 0, 77, term_cb, play_cb, None,
                                       visits, wins, C.byref(best), C.byref(value), C.byref(rootv), C.byref(secs))
        assert rc == 0
        v = np.array(list(visits)); w = np.array(list(wins))
        assert rootv.value == 120_000 and v.sum() == 120_000 and (w <= v).all() and (v > 0).all()
        assert len(streams) == n_searchers              # disjoint random streams per searcher
        assert good >> best.value & 1
        share[n_searchers] = v / v.sum()
    inner = [c for c in range(25) if good >> c & 1]
    assert share[1][inner].sum() > 0.6 and share[4][inner].sum() > 0.6
    assert abs(share[1][inner].sum() - share[4][inner].sum()) < 0.2
