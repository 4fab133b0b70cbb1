"""ctypes binding of oracle/libsquava_oracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "libsquava_oracle.so")

NO_MOVE = -(2 ** 31)
BOARD = np.dtype([("x", "<u4"), ("o", "<u4")])


class Board(C.Structure):
    _fields_ = [("x", C.c_uint32), ("o", C.c_uint32)]


_lib = None


def build():
    src = os.path.join(ODIR, "squava_oracle.c")
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ODIR, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _Oracle(_lib)
    build()
    L = C.CDLL(SO)
    i32p = C.POINTER(C.c_int32)
    L.sqo_mcts_terminal.argtypes = [Board]
    L.sqo_mcts_find_winner.argtypes = [Board]
    L.sqo_mcts_get_result.argtypes = [Board, C.c_int]
    L.sqo_mcts_get_result.restype = C.c_double
    L.sqo_ab_find_winner.argtypes = [Board]
    L.sqo_flags.argtypes = [Board]
    L.sqo_flags.restype = C.c_uint
    L.sqo_classify.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sqo_classify.restype = None
    L.sqo_playouts.argtypes = [Board, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.sqo_playouts.restype = None
    L.sqo_playouts_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64,
                                     C.c_int, C.c_void_p]
    L.sqo_playouts_batch.restype = None
    L.sqo_playout_exact.argtypes = [Board, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sqo_playout_exact.restype = C.c_int64
    for name in ("sqo_ab1_root", "sqo_ab2_root", "sqo_ab5_root", "sqo_ab6_root"):
        f = getattr(L, name)
        f.argtypes = [Board, C.c_int, i32p, i32p, C.POINTER(C.c_uint32), i32p, C.POINTER(C.c_int64)]
        f.restype = None
    L.sqo_negascout_root.argtypes = [Board, C.c_int, i32p, i32p, i32p, C.POINTER(C.c_uint32), i32p, i32p,
                                     C.POINTER(C.c_int64)]
    L.sqo_negascout_root.restype = None
    L.sqo_ab_subtree.argtypes = [C.c_int, Board, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int,
                                 i32p, C.POINTER(C.c_int64)]
    L.sqo_ab_subtree.restype = C.c_int32
    L.sqo_opening3_subtrees.argtypes = [i32p, i32p]
    L.sqo_last_evals.restype = C.c_int64
    _lib = L
    return _Oracle(L)


class _Oracle:
    def __init__(self, L):
        self.L = L

    # ---- tables -----------------------------------------------------------
    def _ints(self, fn, shape, *pre):
        n = int(np.prod(shape))
        buf = (C.c_int * n)()
        cnt = getattr(self.L, fn)(*pre, buf)
        return np.array(buf, dtype=np.int64).reshape(shape), cnt

    def tables(self):
        t = {}
        t["important"] = self._ints("sqo_tables_mcts_important", (9,))[0].tolist()
        t["mcts_quads"], t["mcts_triplets"] = [], []
        for c in range(25):
            a, n = self._ints("sqo_tables_mcts_quads", (6, 4), c)
            t["mcts_quads"].append(a[:n].tolist())
            a, n = self._ints("sqo_tables_mcts_triplets", (12, 3), c)
            t["mcts_triplets"].append(a[:n].tolist())
        t["ab_quads"] = self._ints("sqo_tables_ab_quads", (28, 4))[0].tolist()
        t["ab_triplets"] = self._ints("sqo_tables_ab_triplets", (48, 3))[0].tolist()
        t["ordered"] = self._ints("sqo_tables_ordered_cells", (25,))[0].tolist()
        t["scores_ab"] = self._ints("sqo_tables_scores", (25,), 0)[0].tolist()
        t["scores_opening"] = self._ints("sqo_tables_scores", (25,), 1)[0].tolist()
        t["no2"] = self._ints("sqo_tables_no2", (4, 3))[0].tolist()
        t["no_middle2"] = self._ints("sqo_tables_no_middle2", (4, 4))[0].tolist()
        t["matrices"] = self._ints("sqo_tables_matrices", (8, 25))[0].tolist()
        t["first_moves"] = self._ints("sqo_tables_first_moves", (6,))[0].tolist()
        return t

    # ---- board primitives ---------------------------------------------------
    def terminal(self, x, o):
        return self.L.sqo_mcts_terminal(Board(x, o))

    def mcts_find_winner(self, x, o):
        return self.L.sqo_mcts_find_winner(Board(x, o))

    def get_result(self, x, o, pjm):
        return self.L.sqo_mcts_get_result(Board(x, o), pjm)

    def ab_find_winner(self, x, o):
        return self.L.sqo_ab_find_winner(Board(x, o))

    def flags(self, x, o):
        return self.L.sqo_flags(Board(x, o))

    def classify(self, boards):
        boards = np.ascontiguousarray(boards, dtype=BOARD)
        n = boards.shape[0]
        f = np.empty(n, np.uint8); wm = np.empty(n, np.int8); wa = np.empty(n, np.int8)
        self.L.sqo_classify(boards.ctypes.data, n, f.ctypes.data, wm.ctypes.data, wa.ctypes.data)
        return f, wm, wa

    # ---- playouts -----------------------------------------------------------
    def playouts(self, x, o, to_move, n, seed):
        out = (C.c_uint64 * 4)()
        self.L.sqo_playouts(Board(x, o), to_move, n, seed, out)
        return list(out)

    def playouts_batch(self, boards, to_move, per_leaf, seed, n_threads):
        boards = np.ascontiguousarray(boards, dtype=BOARD)
        tm = np.ascontiguousarray(to_move, dtype=np.int8)
        out = np.zeros((boards.shape[0], 4), np.uint64)
        self.L.sqo_playouts_batch(boards.ctypes.data, tm.ctypes.data, boards.shape[0], per_leaf,
                                  seed, n_threads, out.ctypes.data)
        return out

    def playout_exact(self, x, o, to_move):
        p = (C.c_double * 3)(); mp = C.c_double()
        n = self.L.sqo_playout_exact(Board(x, o), to_move, p, C.byref(mp))
        assert n >= 0
        return list(p), mp.value

    # ---- alpha-beta -----------------------------------------------------------
    def _root(self, fn, x, o, max_depth, scores):
        sc = (C.c_int32 * 25)(*scores); vals = (C.c_int32 * 25)()
        bm = C.c_uint32(); bv = C.c_int32(); lv = C.c_int64()
        getattr(self.L, fn)(Board(x, o), max_depth, sc, vals, C.byref(bm), C.byref(bv), C.byref(lv))
        return list(vals), bm.value, bv.value, lv.value

    def ab1_root(self, x, o, max_depth, scores):
        return self._root("sqo_ab1_root", x, o, max_depth, scores)

    def ab5_root(self, x, o, max_depth, scores):
        return self._root("sqo_ab5_root", x, o, max_depth, scores)

    def ab6_root(self, x, o, max_depth, scores):
        return self._root("sqo_ab6_root", x, o, max_depth, scores)

    def ab2_root(self, x, o, max_depth, scores):
        return self._root("sqo_ab2_root", x, o, max_depth, scores)

    def negascout_root(self, x, o, max_depth, scores):
        """-> (values[25], visit order, best_mask, best_value, first_best, leaves)"""
        sc = (C.c_int32 * 25)(*scores); vals = (C.c_int32 * 25)(); order = (C.c_int32 * 25)()
        bm = C.c_uint32(); bv = C.c_int32(); fb = C.c_int32(); lv = C.c_int64()
        self.L.sqo_negascout_root(Board(x, o), max_depth, sc, vals, order, C.byref(bm), C.byref(bv), C.byref(fb),
                                  C.byref(lv))
        return list(vals), [c for c in order if c >= 0], bm.value, bv.value, fb.value, lv.value

    def ab_subtree(self, dialect, x, o, last_cell, ply, side, carried, max_depth, scores):
        sc = (C.c_int32 * 25)(*scores); lv = C.c_int64()
        v = self.L.sqo_ab_subtree(dialect, Board(x, o), last_cell, ply, side, carried, max_depth,
                                  sc, C.byref(lv))
        return v, lv.value

    def last_evals(self):
        return int(self.L.sqo_last_evals())

    def opening3_subtrees(self):
        f = (C.c_int32 * 150)(); r = (C.c_int32 * 150)()
        n = self.L.sqo_opening3_subtrees(f, r)
        return [(f[i], r[i]) for i in range(n)]
