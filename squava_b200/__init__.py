"""squava_b200 -- B200 (sm_100a) implementation of squava's data-parallel hot path.

This package is a thin ctypes binding over ``libsquava_b200.so`` (C ABI in
``include/squava_b200.h``): random playouts + the terminal-position test, and
the fixed-depth alpha-beta batch.  It is what the tests and ``bench.py`` call;
the reference-facing host programs (``squavam``, ``squavathr``, ``opening3`` ...)
are the C++ sources under ``squava_b200/host`` and bind the same ABI.

There is no CPU fallback: importing works anywhere (so that the symbol table
can be checked on a CPU box), but every compute call needs ``init()`` to have
found an sm_100 device, and raises ``SquavaError`` otherwise.
"""
import ctypes as C
import os

import numpy as np

__all__ = [
    "SquavaError", "BOARD", "SUBTREE", "NO_MOVE", "lib", "lib_path", "init", "shutdown",
    "is_initialised", "classify", "playouts", "playouts_packed", "playouts_dev", "classify_dev",
    "alphabeta_root", "alphabeta_subtrees", "negascout_root", "int32_peak", "launch_count",
    "last_playout_kernel_ms", "debug_philox", "device_count", "sm_count", "EXPORTS",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
# SQUAVA_B200_LIB: another build of the same library (kernel tuning variants, tools/playout_variants.sh)
lib_path = os.environ.get("SQUAVA_B200_LIB") or os.path.join(_HERE, "libsquava_b200.so")

BOARD = np.dtype([("x", "<u4"), ("o", "<u4")])
SUBTREE = np.dtype([("x", "<u4"), ("o", "<u4"), ("last_cell", "i1"), ("ply", "i1"),
                    ("side_to_move", "i1"), ("reserved", "i1"), ("carried", "<i4")])
NO_MOVE = -(2 ** 31)

AB_ALPHABETA, AB_SQUAVATHR, AB_OPENING3, AB_OPENING2, AB_ALPHABETA_AVOID, AB_ABBOOK = 1, 2, 3, 4, 5, 6

# every symbol include/squava_b200.h declares
EXPORTS = [
    "sq_init", "sq_shutdown", "sq_strerror", "sq_last_error", "sq_abi_version", "sq_device_count",
    "sq_sm_count", "sq_classify", "sq_playouts", "sq_alphabeta_root", "sq_alphabeta_subtrees", "sq_negascout_root",
    "sq_classify_dev", "sq_playouts_dev", "sq_int32_peak", "sq_launch_count",
    "sq_last_playout_kernel_ms", "sq_debug_philox", "sq_playouts_packed", "sq_last_error_copy", "sq_playouts_on", "sq_playouts_packed_on", "sq_abi_checked", "sq_debug_check_failures",
    "sq_device_id", "sq_mcts_decide", "sq_mcts_last_error",
]


class MctsResult(C.Structure):
    """sq_mcts_result, include/squava_b200.h"""
    _fields_ = [("n_children", C.c_int32), ("child_move", C.c_int32 * 25), ("child_visits", C.c_double * 25),
                ("child_wins", C.c_double * 25), ("root_visits", C.c_double), ("iterations", C.c_int64), ("batches", C.c_int64),
                ("nodes", C.c_uint64), ("arena_slots", C.c_uint64), ("arena_full", C.c_int32), ("device_ms", C.c_double)]


class SquavaError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        detail = ""
        try:
            detail = lib().sq_last_error().decode()
            name = lib().sq_strerror(status).decode()
        except Exception:  # pragma: no cover
            name = str(status)
        super().__init__("%s: %s (%d)%s" % (where, name, status, (": " + detail) if detail else ""))


_lib = None


def lib():
    """The loaded shared library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise ImportError(
                "libsquava_b200.so is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C squava_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(lib_path)
        vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
        L.sq_init.argtypes = [C.POINTER(C.c_int), i32, u64]
        L.sq_shutdown.restype = None
        L.sq_strerror.argtypes = [i32]; L.sq_strerror.restype = C.c_char_p
        L.sq_last_error.restype = C.c_char_p
        L.sq_sm_count.argtypes = [i32]
        L.sq_classify.argtypes = [vp, i64, vp, vp, vp]
        L.sq_playouts.argtypes = [vp, vp, i64, u64, u64, vp]
        L.sq_playouts_packed.argtypes = [vp, vp, i64, u64, vp]
        L.sq_last_error_copy.argtypes = [C.c_char_p, i32]
        L.sq_debug_check_failures.argtypes = [vp]
        L.sq_playouts_on.argtypes = [i32, vp, vp, i64, u64, u64, vp]
        L.sq_playouts_packed_on.argtypes = [i32, vp, vp, i64, u64, vp]
        L.sq_alphabeta_root.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, vp]
        L.sq_alphabeta_subtrees.argtypes = [vp, i64, i32, i32, vp, vp, vp]
        L.sq_negascout_root.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]
        L.sq_classify_dev.argtypes = [i32, vp, i64, vp, vp, vp, vp]
        L.sq_playouts_dev.argtypes = [i32, vp, vp, i64, u64, u64, u64, vp, vp]
        L.sq_int32_peak.argtypes = [i32, i32, i32, C.POINTER(C.c_double)]
        L.sq_launch_count.restype = u64
        L.sq_last_playout_kernel_ms.argtypes = [i32, C.POINTER(C.c_float)]
        L.sq_debug_philox.argtypes = [vp, vp, vp]
        L.sq_device_id.argtypes = [i32]
        L.sq_mcts_decide.argtypes = [i32, vp, i32, i64, C.c_double, i32, i64, i32, i32, i32, u64, u64, u64, C.POINTER(MctsResult)]
        L.sq_mcts_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _ck(status, where):
    if status != 0:
        raise SquavaError(status, where)


def init(devices=None, seed=0x5155415641):
    """sq_init.  devices: list of CUDA device ordinals (default [0])."""
    L = lib()
    if devices is None:
        _ck(L.sq_init(None, 0, seed), "sq_init")
    else:
        arr = (C.c_int * len(devices))(*devices)
        _ck(L.sq_init(arr, len(devices), seed), "sq_init")


def shutdown():
    lib().sq_shutdown()


def is_initialised():
    return lib().sq_device_count() > 0


def device_count():
    return lib().sq_device_count()


def sm_count(slot=0):
    return lib().sq_sm_count(slot)


def as_boards(boards):
    """Accepts a BOARD array or an (n, 2) integer array of (x, o) masks."""
    b = np.asarray(boards)
    if b.dtype != BOARD:
        b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, 2).view(BOARD).reshape(-1)
    return np.ascontiguousarray(b)


def classify(boards):
    """-> (flags u8[n], winner_mcts_order i8[n], winner_ab_order i8[n])"""
    b = as_boards(boards)
    n = b.shape[0]
    f = np.empty(n, np.uint8); wm = np.empty(n, np.int8); wa = np.empty(n, np.int8)
    _ck(lib().sq_classify(b.ctypes.data, n, f.ctypes.data, wm.ctypes.data, wa.ctypes.data), "sq_classify")
    return f, wm, wa


def playouts(boards, to_move, playouts_per_leaf, stream_id=0, out=None, slot=None):
    """-> u64[n, 4] = {MAX wins, MIN wins, draws, plies} per leaf (host buffers).  slot: run on that one
    device instead of sharding over all (sq_playouts_on)."""
    b = as_boards(boards)
    tm = np.ascontiguousarray(to_move, dtype=np.int8)
    n = b.shape[0]
    if tm.shape[0] != n:
        raise ValueError("to_move length")
    if out is None:
        out = np.empty((n, 4), np.uint64)
    if slot is None:
        _ck(lib().sq_playouts(b.ctypes.data, tm.ctypes.data, n, int(playouts_per_leaf), int(stream_id),
                              out.ctypes.data), "sq_playouts")
    else:
        _ck(lib().sq_playouts_on(int(slot), b.ctypes.data, tm.ctypes.data, n, int(playouts_per_leaf), int(stream_id),
                                 out.ctypes.data), "sq_playouts_on")
    return out


def playouts_packed(boards, to_move, stream_id=0):
    """One playout per leaf -> u8[n]: bits 0-1 outcome (1 MAX won, 2 MIN won, 3 draw), bits 2-7 plies."""
    b = as_boards(boards)
    tm = np.ascontiguousarray(to_move, dtype=np.int8)
    n = b.shape[0]
    if tm.shape[0] != n:
        raise ValueError("to_move length")
    out = np.empty(n, np.uint8)
    _ck(lib().sq_playouts_packed(b.ctypes.data, tm.ctypes.data, n, int(stream_id), out.ctypes.data), "sq_playouts_packed")
    return out


def _stream_ptr(stream):
    if stream is None:
        return None
    return int(getattr(stream, "cuda_stream", stream))


def classify_dev(boards_ptr, n, flags_ptr, wm_ptr, wa_ptr, slot=0, stream=None):
    _ck(lib().sq_classify_dev(slot, boards_ptr, n, flags_ptr, wm_ptr, wa_ptr, _stream_ptr(stream)),
        "sq_classify_dev")


def playouts_dev(leaves_ptr, to_move_ptr, n_leaves, playouts_per_leaf, out_ptr, leaf_index_base=0,
                 stream_id=0, slot=0, stream=None):
    """Device-resident form: raw device pointers (e.g. torch ``tensor.data_ptr()``); asynchronous."""
    _ck(lib().sq_playouts_dev(slot, leaves_ptr, to_move_ptr, n_leaves, leaf_index_base,
                              int(playouts_per_leaf), int(stream_id), out_ptr, _stream_ptr(stream)),
        "sq_playouts_dev")


def _scores(scores):
    s = np.ascontiguousarray(scores, dtype=np.int8).reshape(-1)
    if s.shape[0] != 25:
        raise ValueError("scores must have 25 entries")
    return s


def alphabeta_root(positions, dialect, max_depth, scores):
    """-> (values i32[n,25], best_value i32[n], best_mask u32[n], nodes u64[n])"""
    b = as_boards(positions)
    n = b.shape[0]
    s = _scores(scores)
    values = np.empty((n, 25), np.int32); bv = np.empty(n, np.int32)
    bm = np.empty(n, np.uint32); nodes = np.empty(n, np.uint64)
    _ck(lib().sq_alphabeta_root(b.ctypes.data, n, dialect, max_depth, s.ctypes.data, values.ctypes.data,
                                bv.ctypes.data, bm.ctypes.data, nodes.ctypes.data), "sq_alphabeta_root")
    return values, bv, bm, nodes


def negascout_root(positions, max_depth, scores):
    """-> dict(values i32[n,25], visit_order i8[n,25], best_value, best_mask, first_best, nodes)"""
    b = as_boards(positions)
    n = b.shape[0]
    s = _scores(scores)
    r = dict(values=np.empty((n, 25), np.int32), visit_order=np.empty((n, 25), np.int8),
             best_value=np.empty(n, np.int32), best_mask=np.empty(n, np.uint32), first_best=np.empty(n, np.int8),
             nodes=np.empty(n, np.uint64))
    _ck(lib().sq_negascout_root(b.ctypes.data, n, max_depth, s.ctypes.data, r["values"].ctypes.data,
                                r["visit_order"].ctypes.data, r["best_value"].ctypes.data, r["best_mask"].ctypes.data,
                                r["first_best"].ctypes.data, r["nodes"].ctypes.data), "sq_negascout_root")
    return r


def mcts_decide(x, o, player_just_moved, iterations, uctk=1.0, ucb_factor2=True, batch=65536, playouts_per_leaf=1, walkers=0,
                sync_levels=-1, max_nodes=0, stream_base=0, seed=1, slot=0):
    """sq_mcts_decide: one UCT decision with the tree in HBM -> dict(moves, visits, wins, root_visits, best_move, ...).
    best_move is argmax UCB1 over the root's children (mcts.go:196-198)."""
    b = np.array([(x, o)], dtype=BOARD)
    r = MctsResult()
    st = lib().sq_mcts_decide(slot, b.ctypes.data, player_just_moved, iterations, uctk, 1 if ucb_factor2 else 0, batch, playouts_per_leaf,
                              walkers, sync_levels, max_nodes, stream_base, seed, C.byref(r))
    if st != 0:
        raise RuntimeError("sq_mcts_decide: %d %s" % (st, lib().sq_mcts_last_error().decode()))
    n = r.n_children
    moves = [r.child_move[i] for i in range(n)]
    visits = np.array([r.child_visits[i] for i in range(n)]); wins = np.array([r.child_wins[i] for i in range(n)])
    best = None
    if n:
        lg = (2.0 if ucb_factor2 else 1.0) * np.log(max(r.root_visits, 1.0))
        score = wins / (visits + 1e-6) + uctk * np.sqrt(lg / (visits + 1e-6))
        best = moves[int(np.argmax(score))]
    return dict(moves=moves, visits=visits, wins=wins, root_visits=r.root_visits, best_move=best, iterations=r.iterations,
                batches=r.batches, nodes=r.nodes, arena_slots=r.arena_slots, arena_full=bool(r.arena_full), device_ms=r.device_ms)


def alphabeta_subtrees(subtrees, dialect, max_depth, scores):
    """subtrees: SUBTREE array -> (value i32[n], nodes u64[n])"""
    t = np.ascontiguousarray(subtrees, dtype=SUBTREE)
    n = t.shape[0]
    s = _scores(scores)
    value = np.empty(n, np.int32); nodes = np.empty(n, np.uint64)
    _ck(lib().sq_alphabeta_subtrees(t.ctypes.data, n, dialect, max_depth, s.ctypes.data,
                                    value.ctypes.data, nodes.ctypes.data), "sq_alphabeta_subtrees")
    return value, nodes


def check_failures():
    """Checked build: (failed in-kernel checks since the last call, code of the first, its two arguments)."""
    out = np.zeros(4, np.uint32)
    _ck(lib().sq_debug_check_failures(out.ctypes.data), "sq_debug_check_failures")
    return tuple(int(v) for v in out)


def int32_peak(kind=2, reps=5, slot=0):
    """Measured INT32 thread-ops/s: kind 0 LOP3, 1 IMAD, 2 both pipes interleaved."""
    v = C.c_double()
    _ck(lib().sq_int32_peak(slot, kind, reps, C.byref(v)), "sq_int32_peak")
    return v.value


def launch_count():
    return int(lib().sq_launch_count())


def last_playout_kernel_ms(slot=0):
    v = C.c_float()
    _ck(lib().sq_last_playout_kernel_ms(slot, C.byref(v)), "sq_last_playout_kernel_ms")
    return v.value


def debug_philox(counter, key):
    c = np.ascontiguousarray(counter, dtype=np.uint32); k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, np.uint32)
    _ck(lib().sq_debug_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data), "sq_debug_philox")
    return out
