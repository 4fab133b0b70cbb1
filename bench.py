#!/usr/bin/env python3
"""bench.py -- random playouts/s on BASELINE.json's configs[1]:
"batched random playouts from 4096 synthetic mid-game positions (8-14 marks), 10^6
playouts each, 1 GPU"; with --gpus N every rank runs its own 4096-leaf shard (weak
scaling, leaf index keyed globally) and the per-leaf count vectors are merged with one
NCCL all-reduce per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  Keys follow the driver's contract; see DESIGN.md
"Measurement" for how each number is obtained.

  value      device-resident throughput: leaves already in HBM, sq_playouts_dev on the
             current torch stream, CUDA events, max over ranks.
  e2e        the reference-facing call (sq_playouts, host buffers): pinned-host -> device
             copy of the leaves and device -> host read of the counts inside the timed
             region, every step.
  roofline   INT32 issue roofline: algorithmic ops (64 per ply + 16 per playout, SURVEY
             8d) / kernel time / measured INT32 issue rate (sq_int32_peak, both pipes).
  cpu_baseline  the CPU oracle (a C port of the Go reference's table-walk playout) on
             all host threads, bounded sample.  Rank 0, N=1 only.

  secondary  the other BASELINE configs, each timed once and parity-checked inside the run
             (N=1: c0 squavam decision, c2_1e8 host tree, c3 alpha-beta batch; every N: c4 opening3
             sharded over the ranks with an NCCL merge, and `strong`: the SAME 4096 leaves split
             over the ranks).  --no-secondary skips them.

--impl reference times that same CPU port as the reference arm (the Go reference cannot
be built here: no Go toolchain).
"""
import argparse
import ctypes
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LEAVES = 4096
PER_LEAF = 1_000_000
METRIC = "random playouts/sec"
UNIT = "playouts/s"
WORKLOAD = "configs[1]: 4096 synthetic mid-game positions (8-14 marks) x 1e6 random playouts each"
OPS_PER_PLY, OPS_PER_PLAYOUT = 64, 16          # the agreed algorithmic op count


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_port_rate(boards, to_move, target_s, threads, seed=1):
    """Times the CPU oracle (C port of the reference's playout) on a bounded sample.
    -> (playouts/s, per_leaf used, seconds)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.load()
    t0 = time.perf_counter()
    o.playouts_batch(boards, to_move, 50, seed, threads)
    probe = time.perf_counter() - t0
    rate = boards.shape[0] * 50 / max(probe, 1e-6)
    per_leaf = int(max(50, min(200_000, rate * target_s / boards.shape[0])))
    t0 = time.perf_counter()
    out = o.playouts_batch(boards, to_move, per_leaf, seed + 1, threads)
    dt = time.perf_counter() - t0
    assert int(out[:, :3].sum()) == boards.shape[0] * per_leaf
    return boards.shape[0] * per_leaf / dt, per_leaf, dt


def source_hash():
    """sha256 of the sources the playout kernel is compiled from: a committed ncu count is only used
    while it describes THIS kernel (profiles/playout_kernel_traffic.json carries the hash it was taken at)."""
    h = hashlib.sha256()
    for f in ("kernels.cuh", "sq_device.cuh"):
        h.update(open(os.path.join(ROOT, "squava_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def oracle_handle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib.load()


def lpt_assign(weights, ranks):
    """largest first, each to the least loaded rank -> list of index lists"""
    order = sorted(range(len(weights)), key=lambda i: -weights[i])
    load, mine = [0.0] * ranks, [[] for _ in range(ranks)]
    for i in order:
        k = min(range(ranks), key=lambda r: load[r])
        mine[k].append(i); load[k] += weights[i]
    return mine


SCORES_AB = [3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3]          # alphabeta.go:343-349
SCORES_OPENING = [3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 2, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3]     # opening3.go:366-372
AB_OPS_PER_EVAL = 48


def secondary_c3(sq, peak_ops):
    """configs[3]: 1024 positions (8/10/12/14 marks), src/alphabeta depth 10, one sq_alphabeta_root call with host
    buffers.  Parity: the CPU oracle searches a bounded sample (smallest trees first, all host threads, ~6 s) and
    every sampled row + tie mask is compared; the eight-mark trees are covered by tests/golden/c3_*.json in the
    test suite.  Rate: reference-order evaluator calls (the oracle's sequential count for the whole batch, measured
    once and committed with the hash of the positions) per second."""
    from concurrent.futures import ThreadPoolExecutor
    from squava_b200 import positions
    boards, _ = positions.midgame_batch(1024, marks=(8, 10, 12, 14), seed=0xC4)
    sq.alphabeta_root(boards[:8], 1, 4, SCORES_AB)                               # warm-up
    best, l0 = None, sq.launch_count()
    for rep in range(3):
        t0 = time.perf_counter()
        values, bv, bm, nodes = sq.alphabeta_root(boards, 1, 10, SCORES_AB)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    launches = (sq.launch_count() - l0) // 3
    rec = {"workload": "configs[3]: 1024 positions (8/10/12/14 marks), AB-1 depth 10, sq_alphabeta_root (host buffers)",
           "seconds": best, "gpu_evaluator_calls": int(nodes.sum()), "gpu_launches": launches}
    ref = None
    try:
        ref = json.load(open(os.path.join(ROOT, "tests", "golden", "c3_reference_order_counts.json")))
        if ref.get("positions_sha256") != hashlib.sha256(boards.tobytes()).hexdigest():
            ref = None
    except Exception:
        ref = None
    if ref:
        evals = ref["oracle_evaluator_calls"]
        rec["ref_order_evals_per_s"] = evals / best
        rec["roofline"] = {"bound": "int32_issue", "ops_per_eval": AB_OPS_PER_EVAL, "achieved": AB_OPS_PER_EVAL * evals / best / 1e12,
                           "peak": peak_ops / 1e12, "unit": "Tops/s", "frac": AB_OPS_PER_EVAL * evals / best / peak_ops,
                           "frac_gpu_evals": AB_OPS_PER_EVAL * int(nodes.sum()) / best / peak_ops,
                           "note": "reference-order evaluator calls (oracle's sequential count, %s) x 48 ops" % ref["source"]}
    # bounded oracle sample: positions with the most marks first (smallest trees), until ~6 s of wall time
    o = oracle_handle()
    order = sorted(range(1024), key=lambda i: -bin(int(boards[i]["x"]) | int(boards[i]["o"])).count("1"))
    threads = host_threads()
    t0 = time.perf_counter()
    done, evals_cpu, mism = 0, 0, 0

    def one(i):
        vals, obm, obv, _ = o.ab1_root(int(boards[i]["x"]), int(boards[i]["o"]), 10, SCORES_AB)
        return i, vals, obm, obv, o.last_evals()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        pos = 0
        while pos < len(order) and time.perf_counter() - t0 < 6.0:
            chunk = order[pos:pos + 4 * threads]
            pos += len(chunk)
            for i, vals, obm, obv, ev in ex.map(one, chunk):
                done += 1; evals_cpu += ev
                if values[i].tolist() != vals or int(bm[i]) != obm or int(bv[i]) != obv:
                    mism += 1
    cpu_s = time.perf_counter() - t0
    rec["mismatching_positions"] = mism
    rec["parity_sample"] = "%d of 1024 positions (most marks first) searched by the CPU oracle in this run" % done
    rec["cpu_baseline"] = {"value": evals_cpu / cpu_s, "unit": "evaluator calls/s", "cores": threads, "kind": "port",
                           "sample": "%d positions, %.1f s" % (done, cpu_s)}
    return rec


def secondary_negascout(sq):
    """playoff5's N engine (src/negascout) as a batch and as a single decision: sq_negascout_root with host buffers, split
    over many warps (the default) beside one warp per position (SQ_NS_SPLIT=0, round 1's schedule).  The search is
    order-dependent, so parity is everything: values, visiting order, movekeeper set and leaf counter of the two
    schedules must be identical here, and a sample is compared with the CPU oracle."""
    from squava_b200 import positions
    rec = {}
    o = oracle_handle()
    for name, count, marks, depth, sample in [("batch", 1024, 8, 8, 2), ("one_position", 1, 2, 6, 0)]:
        boards, _ = positions.midgame_batch(count, marks=(marks,), seed=4000 + marks)
        sq.negascout_root(boards, depth, SCORES_AB)                              # warm-up: the flights' pinned buffers are made on first use
        split_s = None
        for rep in range(2):
            l0 = sq.launch_count()
            t0 = time.perf_counter(); split = sq.negascout_root(boards, depth, SCORES_AB); dt = time.perf_counter() - t0
            split_s = dt if split_s is None else min(split_s, dt)
            launches = sq.launch_count() - l0
        os.environ["SQ_NS_SPLIT"] = "0"
        try:
            t0 = time.perf_counter(); plain = sq.negascout_root(boards, depth, SCORES_AB); plain_s = time.perf_counter() - t0
        finally:
            del os.environ["SQ_NS_SPLIT"]
        same = all((split[k] == plain[k]).all() for k in ("values", "visit_order", "best_mask", "best_value", "first_best", "nodes"))
        mism = 0
        order = sorted(range(count), key=lambda i: int(split["nodes"][i]))[:sample]
        for i in order:
            vals, vis, bm, bv, fb, leaves = o.negascout_root(int(boards[i]["x"]), int(boards[i]["o"]), depth, SCORES_AB)
            ok = (split["values"][i].tolist() == vals and [int(c) for c in split["visit_order"][i] if c >= 0] == vis
                  and (int(split["best_mask"][i]), int(split["best_value"][i]), int(split["first_best"][i])) == (bm, bv, fb)
                  and int(split["nodes"][i]) == leaves)
            mism += 0 if ok else 1
        rec[name] = {"workload": "%d position(s), %d marks, negascout depth %d, sq_negascout_root (host buffers)" % (count, marks, depth),
                     "seconds": split_s, "seconds_one_warp_per_position": plain_s, "static_value_calls": int(split["nodes"].sum()),
                     "calls_per_s": int(split["nodes"].sum()) / split_s, "gpu_launches": launches,
                     "schedules_identical": bool(same), "oracle_sample": len(order), "oracle_mismatches": mism}
    return rec


def opening3_subtrees(sq):
    """opening3.go:66-128: the (first move, reply) pairs not equivalent under the program's 8 symmetries
    (opening3.go:379-451) to a pair computed earlier: 85 = 14 + 24 + 14 + 14 + 14 + 5."""
    def images(c):
        x, y = divmod(c, 5)
        return [5 * a + b for a, b in ((x, y), (y, 4 - x), (4 - x, 4 - y), (4 - y, x), (4 - x, y), (x, 4 - y), (4 - y, 4 - x), (y, x))]
    done, pairs = set(), []
    for f in (0, 1, 2, 6, 7, 12):                       # firstMoves, opening3.go:456-463
        for r in range(25):
            if r == f or any((a, b) in done for a, b in zip(images(f), images(r))):
                continue
            done.add((f, r)); pairs.append((f, r))
    assert len(pairs) == 85
    st = np.zeros(len(pairs), sq.SUBTREE)
    for i, (f, r) in enumerate(pairs):
        st[i] = (1 << f, 1 << r, r, 2, 1, 0, SCORES_OPENING[f])
    return pairs, st


def secondary_c4(sq, rank, world, dev, dist, torch, peak_ops):
    """configs[4], alpha-beta half: opening3's 85 subtrees (AB-3, -d 10) sharded over the ranks -- by the node counts of
    a depth-7 pilot search (largest subtree first, each to the least loaded rank), every rank searching its share on
    its own GPU -- and ONE NCCL all-reduce of the zero-padded value vector.  Parity: all 85 values against the
    oracle's depth-10 vectors (tests/golden/c4_opening3_depth10.json)."""
    pairs, st = opening3_subtrees(sq)
    sq.alphabeta_subtrees(st[:4], 3, 5, SCORES_OPENING)                          # warm-up
    runs = []
    for rep in range(2):                                                         # the first run also grows the node pool; both are reported
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world > 1:
            _, pilot = sq.alphabeta_subtrees(st, 3, 7, SCORES_OPENING)
            # evaluator counts depend on how the GPU happened to schedule the search: ONE rank's counts are the weights for
            # everybody (a broadcast of 85 integers), so that all ranks derive the same partition
            weights = torch.from_numpy(pilot.astype(np.int64)).to(dev)
            dist.broadcast(weights, src=0)
            mine = lpt_assign([float(v) for v in weights.cpu().numpy()], world)[rank]
        else:
            mine = list(range(len(pairs)))                                           # one rank: nothing to partition, no pilot search
        vals = torch.zeros(len(pairs), dtype=torch.int64, device=dev)
        nodes = torch.zeros(len(pairs), dtype=torch.int64, device=dev)
        if mine:
            v, n = sq.alphabeta_subtrees(st[mine], 3, 10, SCORES_OPENING)
            vals[mine] = torch.from_numpy(v.astype(np.int64)).to(dev)
            nodes[mine] = torch.from_numpy(n.astype(np.int64)).to(dev)
        if world > 1:
            dist.all_reduce(vals); dist.all_reduce(nodes)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        runs.append(float(t.item()))
    t = torch.tensor([min(runs)], dtype=torch.float64, device=dev)
    secs = float(t.item())
    got = vals.cpu().numpy().tolist()
    rec = {"workload": "configs[4] alpha-beta: opening3's 85 subtrees, AB-3 depth 10, sharded over %d rank(s) by pilot-search size (pilot inside the timed region when there is more than one rank), NCCL merge" % world,
           "seconds": secs, "seconds_runs": runs, "gpu_evaluator_calls": int(nodes.sum().item()), "subtrees_on_rank0": len(mine) if rank == 0 else None}
    gp = os.path.join(ROOT, "tests", "golden", "c4_opening3_depth10.json")
    if os.path.exists(gp):
        g = json.load(open(gp))
        want = {(c["first"], c["reply"]): c["value"] for c in g["cases"]}
        rec["mismatching_subtrees"] = sum(1 for (f, r), v in zip(pairs, got) if want[(f, r)] != v)
        rec["ref_order_evals_per_s"] = g["total_evals"] / secs
        rec["roofline"] = {"bound": "int32_issue", "ops_per_eval": AB_OPS_PER_EVAL, "achieved": AB_OPS_PER_EVAL * g["total_evals"] / secs / 1e12,
                           "peak": peak_ops * world / 1e12, "unit": "Tops/s", "frac": AB_OPS_PER_EVAL * g["total_evals"] / secs / (peak_ops * world)}
        rec["cpu_baseline"] = {"value": g["total_evals"] / g["wall_seconds"], "unit": "evaluator calls/s", "cores": g["procs"], "kind": "port",
                               "sample": "all 85 subtrees, measured when the golden vectors were made (%.0f s; not on this box)" % g["wall_seconds"]}
    else:
        rec["mismatching_subtrees"] = None
    return rec


def secondary_c4_mcts(sq, rank, world, dev, dist, torch, playouts_per_position):
    """configs[4], MCTS half: root-parallel MCTS from opening3's 85 two-mark positions.  Every rank grows its OWN tree
    per position on its own GPU (C++ arena tree of the host library, leaf batches of 1024 leaves x 65536 playouts,
    its own host RNG seed and random-stream ids), then the per-root-move {visits, wins} vectors of all positions --
    85 x 2 x 25 doubles -- are merged with ONE NCCL all-reduce and the move is chosen from the merged statistics
    (argmax UCB1, mcts.go:196-198).  Checked inside the run: every iteration is counted once."""
    C = ctypes
    L = host_lib()
    L.sqh_root_parallel_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]
    pairs, _ = opening3_subtrees(sq)
    ppl, batch = 65536, 1024
    iters = int(playouts_per_position)
    stats = np.zeros((len(pairs), 2, 25), np.float64)
    visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
    best, value, rootv, secs = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, (f, r) in enumerate(pairs):
        rc = L.sqh_root_parallel_uct(1 << f, 1 << r, -1, iters, 1.0, 1, batch, ppl, 0, 1, 0, ((rank + 1) << 56) | (i << 44),
                                     1000 + 85 * rank + i, None, None, None, visits, wins, C.byref(best), C.byref(value),
                                     C.byref(rootv), C.byref(secs))
        assert rc == 0 and rootv.value == iters
        stats[i, 0] = list(visits); stats[i, 1] = list(wins)
    t_stats = torch.from_numpy(stats).to(dev)
    if world > 1:
        dist.all_reduce(t_stats)                                       # the merge: 85 x 2 x 25 doubles
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    merged = t_stats.cpu().numpy()
    assert (merged[:, 0].sum(axis=1) == iters * world).all() and (merged[:, 1] <= merged[:, 0]).all()
    moves = []
    for i in range(len(pairs)):
        v, w = merged[i, 0], merged[i, 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = np.where(v > 0, w / v + np.sqrt(2.0 * np.log(v.sum()) / v), -1.0)
        moves.append(int(np.argmax(u)))
    secs_all = float(t.item())
    return {"workload": "configs[4] MCTS: root-parallel trees from the 85 opening3 positions, %.3g playouts per position and rank (leaf batches of %d x %d), %d rank(s), one NCCL all-reduce of the 85x2x25 root statistics" % (iters, batch, ppl, world),
            "seconds": secs_all, "value": len(pairs) * iters * world / secs_all, "unit": UNIT, "playouts_per_s_per_gpu": len(pairs) * iters / secs_all,
            "chosen_moves": moves}


def host_lib():
    L = ctypes.CDLL(os.path.join(ROOT, "squava_b200", "host", "libsquava_host.so"))
    C = ctypes
    L.sqh_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.sqh_fast_uct.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int),
                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    return L


def secondary_c0_c2(sq, iters_c2):
    """configs[0]: one squavam move decision from the empty board, -i 10000 -u 1.0, computer first (UCT with squavam.go's
    UCB1, leaf batches of 256), and configs[2] at 1e8 iterations: ONE host tree, one playout per leaf, GPU leaf batches
    of 65536 (the arena tree of squavam -fast).  Both through the C++ host library that squavam itself is built from.
    Consistency checks inside the run: every iteration is counted once (root visits == iterations, children sum to it)."""
    C = ctypes
    L = host_lib()
    moves = (C.c_int * 25)(); visits = (C.c_double * 25)(); wins = (C.c_double * 25)()
    nch, best, value, leaves = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    out = {}
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        rc = L.sqh_uct(0, 0, 1, 10000, 1.0, 0, 256, 1, None, None, None, None, C.byref(nch), moves, visits, wins,
                       C.byref(best), C.byref(value), C.byref(leaves))
        ts.append(time.perf_counter() - t0)
        assert rc == 0 and leaves.value == 10000 and sum(visits[i] for i in range(nch.value)) == 10000
    out["c0"] = {"workload": "configs[0]: squavam -C -i 10000 -u 1.0, one move decision from the empty board (leaf batches of 256)",
                 "seconds": min(ts), "iterations_per_s": 10000 / min(ts), "best_move": best.value}
    rootv, nodes = C.c_double(), C.c_ulonglong()
    secs = (C.c_double * 3)()
    t0 = time.perf_counter()
    rc = L.sqh_fast_uct(0, 0, 1, iters_c2, 1.0, 0, 65536, 1, 0, None, None, None, None, C.byref(nch), moves, visits, wins,
                        C.byref(best), C.byref(value), C.byref(rootv), secs, C.byref(nodes))
    dt = time.perf_counter() - t0
    assert rc == 0 and rootv.value == iters_c2 and sum(visits[i] for i in range(nch.value)) == iters_c2
    assert all(0 <= wins[i] <= visits[i] for i in range(nch.value))
    out["c2_1e8"] = {"workload": "configs[2] at %.0e iterations: one host tree, one playout per leaf, GPU leaf batches of 65536" % iters_c2,
                     "seconds": dt, "iterations_per_s": iters_c2 / dt, "host_threads": host_threads(),
                     "select_seconds": secs[0], "gpu_seconds": secs[1], "update_seconds": secs[2], "tree_node_slots": int(nodes.value),
                     "best_move": best.value}
    # configs[2] with the tree in HBM (sq_mcts_decide): the same decision, select / playout / update all on the device.
    # 1e8 first; the full 1e9 of BASELINE.json when that took a few seconds.  Checked inside the run: the books.
    for iters in (iters_c2, 10 * iters_c2):
        t0 = time.perf_counter()
        r = sq.mcts_decide(0, 0, 1, iters, uctk=1.0, ucb_factor2=False, batch=65536, playouts_per_leaf=1, seed=iters)
        dt = time.perf_counter() - t0
        assert r["iterations"] == iters and r["root_visits"] == iters and iters - 4096 <= r["visits"].sum() <= iters
        assert (r["wins"] <= r["visits"]).all()
        out["c2_gpu_tree_%.0e" % iters] = {
            "workload": "configs[2] at %.0e iterations: one tree in HBM (sq_mcts_decide), one playout per leaf, batches of 65536 walks" % iters,
            "seconds": dt, "device_seconds": r["device_ms"] * 1e-3, "iterations_per_s": iters / dt, "batches": int(r["batches"]),
            "tree_node_slots": int(r["nodes"]), "arena_slots": int(r["arena_slots"]), "arena_full": r["arena_full"],
            "best_move": r["best_move"], "root_children_visits_max_share": float(r["visits"].max() / r["visits"].sum())}
        if dt > 12.0:
            break
    return out


def run_reference(args):
    """Reference arm: the CPU port of the Go playout loop on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from squava_b200 import positions
    boards, to_move = positions.midgame_batch(N_LEAVES)
    threads = host_threads()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.load()
    t0 = time.perf_counter()
    o.playouts_batch(boards, to_move, 20, 7, threads)
    rate = N_LEAVES * 20 / (time.perf_counter() - t0)
    total_steps = args.steps + args.warmup
    per_step_s = min(8.0, 150.0 / max(1, total_steps))
    per_leaf = int(max(20, min(PER_LEAF, rate * per_step_s / N_LEAVES)))
    for w in range(args.warmup):
        o.playouts_batch(boards, to_move, per_leaf, 100 + w, threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        o.playouts_batch(boards, to_move, per_leaf, 200 + s, threads)
    dt = time.perf_counter() - t0
    value = N_LEAVES * per_leaf * args.steps / dt
    sample = "%d positions x %d playouts per step (bounded sample of the 1e6-per-leaf workload)" % (N_LEAVES, per_leaf)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C port of the Go reference's table-walk playout (oracle/squava_oracle.c); "
                                 "no Go toolchain on this image, so the reference itself cannot be built"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--per-leaf", type=int, default=PER_LEAF, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs (c0, c2_1e8, c3, c4, strong)")
    ap.add_argument("--c2-iterations", type=float, default=1e8, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its
    # "NCCL version ..." banner on stdout): keep the real stdout aside and point fd 1 at stderr.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import squava_b200 as sq
    from squava_b200 import positions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the squava_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sq.init(devices=[local_rank], seed=positions.SEED_C2)
    W = max(3, args.warmup)
    K = args.steps
    per_leaf = args.per_leaf

    # every rank: its own 4096-leaf shard of a global batch of world*4096 leaves
    boards_all, tm_all = positions.midgame_batch(N_LEAVES * world)
    lo = rank * N_LEAVES
    boards, to_move = boards_all[lo:lo + N_LEAVES], tm_all[lo:lo + N_LEAVES]
    d_boards = torch.from_numpy(boards.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dev)
    d_tm = torch.from_numpy(to_move.copy()).to(dev)
    d_merged = torch.zeros((N_LEAVES * world, 4), dtype=torch.int64, device=dev)
    d_out = d_merged[lo:lo + N_LEAVES]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step(i):
        flush.fill_(i & 0xff)                                           # L2 flush between iterations
        if world > 1:
            d_merged.zero_()
        sq.playouts_dev(d_boards.data_ptr(), d_tm.data_ptr(), N_LEAVES, per_leaf, d_out.data_ptr(),
                        leaf_index_base=lo, stream_id=i, stream=stream)
        if world > 1:
            dist.all_reduce(d_merged)                                   # merge the per-leaf count vectors

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sq.launch_count()
    barrier()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        flush.fill_(i & 0xff)
        if world > 1:
            d_merged.zero_()
        k_ev[i][0].record()
        sq.playouts_dev(d_boards.data_ptr(), d_tm.data_ptr(), N_LEAVES, per_leaf, d_out.data_ptr(),
                        leaf_index_base=lo, stream_id=W + i, stream=stream)
        k_ev[i][1].record()
        if world > 1:
            dist.all_reduce(d_merged)
    t_stop.record()
    barrier()
    launches = sq.launch_count() - launches0
    clocks = sampler.finish()
    elapsed_ms = t_start.elapsed_time(t_stop)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    last_kernel_ms = sq.last_playout_kernel_ms()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    res = d_out.cpu().numpy()
    assert (res[:, :3].sum(axis=1) == per_leaf).all(), "every playout must end in exactly one outcome"
    plies_per_step = int(res[:, 3].sum())
    value = N_LEAVES * world * per_leaf * K / (elapsed_ms * 1e-3)

    # ---- end to end through the host-buffer C ABI call ------------------------------
    h_boards = torch.from_numpy(boards.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).pin_memory()
    h_tm = torch.from_numpy(to_move.copy()).pin_memory()
    h_out = torch.zeros((N_LEAVES, 4), dtype=torch.int64).pin_memory()
    L = sq.lib()

    def e2e_step(i):
        rc = L.sq_playouts(h_boards.data_ptr(), h_tm.data_ptr(), N_LEAVES, per_leaf, 1000 + i, h_out.data_ptr())
        if rc:
            raise sq.SquavaError(rc, "sq_playouts")
        if world > 1:                                                   # merge shards, read the result back
            d_merged.zero_()
            d_out.copy_(h_out, non_blocking=True)
            dist.all_reduce(d_merged)
            return int(d_merged[:, :3].sum().item())
        return int(h_out[:, :3].sum().item())

    Ke = max(3, K // 2)
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        got = e2e_step(2 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert got == N_LEAVES * world * per_leaf
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N_LEAVES * world * per_leaf * Ke / float(t.item())

    # ---- roofline of the playout kernel ------------------------------------------------
    peak_mixed = sq.int32_peak(kind=2, reps=5)
    peak_alu = sq.int32_peak(kind=0, reps=3)
    peak_fma = sq.int32_peak(kind=1, reps=3)
    alg_ops = OPS_PER_PLY * plies_per_step + OPS_PER_PLAYOUT * N_LEAVES * per_leaf
    model_rate = alg_ops / (kernel_ms * 1e-3)
    # What the kernel really issues: thread-instructions per ply counted by ncu (smsp__inst_executed of one
    # launch of this workload) -- used only while the committed count was taken from THIS source (hash).
    traffic, inst_per_ply, prof_note = None, None, "no committed ncu count"
    tp = os.path.join(ROOT, "profiles", "playout_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp))
            if prof.get("source_sha256") == source_hash():
                traffic = prof.get("dram_bytes_per_launch")
                inst_per_ply = prof["warp_instructions_per_launch"] * 32.0 / prof["plies_per_launch"]
                prof_note = prof.get("source")
            else:
                prof_note = "committed ncu count is stale (kernel source changed since it was taken): not used"
        except Exception as e:
            prof_note = "unreadable: %s" % e
    if inst_per_ply and per_leaf == PER_LEAF:
        # achieved = executed thread-instructions per second: a true fraction of the issue ceiling (<= 1)
        executed = inst_per_ply * plies_per_step
        achieved, basis = executed / (kernel_ms * 1e-3), "executed thread-instructions (ncu count per ply x plies of this run)"
    else:
        achieved, basis = min(model_rate, peak_mixed), "SURVEY 8d op model, capped at the peak (no fresh ncu count)"
    roofline = {
        "bound": "int32_issue", "achieved": achieved / 1e12, "peak": peak_mixed / 1e12, "unit": "Tops/s",
        "frac": achieved / peak_mixed, "traffic": traffic, "basis": basis,
        "kernel": "sq::playout_kernel", "kernel_ms": kernel_ms, "kernel_ms_last_internal_events": last_kernel_ms,
        "model_ops_per_launch": alg_ops, "ops_model": "64 per ply + 16 per playout (SURVEY.md 8d)",
        "model_tops": model_rate / 1e12, "model_over_peak": model_rate / peak_mixed,
        "model_note": "the kernel issues fewer instructions per ply than the survey's 64+16 budget, so the model rate can exceed the issue peak; frac is the executed-instruction fraction",
        "plies_per_launch": plies_per_step, "peak_source": "sq_int32_peak measured in this run (LOP3+IMAD interleaved)",
        "peak_alu_only_tops": peak_alu / 1e12, "peak_fma_only_tops": peak_fma / 1e12,
        "algorithmic_bytes_per_launch": N_LEAVES * (9 + 32), "ncu_count": prof_note,
    }
    if inst_per_ply:
        roofline["executed_thread_inst_per_ply"] = inst_per_ply
    assert roofline["frac"] <= 1.0 + 1e-9

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "leaves_per_gpu": N_LEAVES, "playouts_per_leaf": per_leaf,
                   "parallelism": "leaf-sharded x%d, one NCCL all-reduce of the [leaves,4] count matrix per step" % world
                   if world > 1 else "1 GPU",
                   "l2": "256 MiB write between iterations (inside the timed region); working set is 164 KB",
                   "mean_plies_per_playout": plies_per_step / (N_LEAVES * per_leaf)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_LEAVES * 9 * world,
                "d2h_bytes_per_step": N_LEAVES * 32 * world, "steps": Ke, "api": "sq_playouts (host buffers, C ABI)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }

    if not args.no_secondary:
        sec = {}
        # ---- strong scaling: the SAME 4096 leaves x 1e6 playouts, split over the ranks, one NCCL merge per step ----
        s_lo, s_hi = N_LEAVES * rank // world, N_LEAVES * (rank + 1) // world
        sb, st_ = boards_all[s_lo:s_hi], tm_all[s_lo:s_hi]
        ds_b = torch.from_numpy(sb.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(dev)
        ds_t = torch.from_numpy(st_.copy()).to(dev)
        ds_m = torch.zeros((N_LEAVES, 4), dtype=torch.int64, device=dev)

        def strong_step(i):
            flush.fill_(i & 0xff)
            if world > 1:
                ds_m.zero_()
            sq.playouts_dev(ds_b.data_ptr(), ds_t.data_ptr(), s_hi - s_lo, per_leaf, ds_m[s_lo:s_hi].data_ptr(),
                            leaf_index_base=s_lo, stream_id=5000 + i, stream=stream)
            if world > 1:
                dist.all_reduce(ds_m)
        Ks = max(3, K // 2)
        for i in range(3):
            strong_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(Ks):
            strong_step(3 + i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        got = ds_m.cpu().numpy()
        assert (got[:, :3].sum(axis=1) == per_leaf).all()
        sec["strong"] = {"workload": "the SAME 4096 leaves x %d playouts split over %d rank(s), NCCL all-reduce of the [4096,4] counts per step" % (per_leaf, world),
                         "value": N_LEAVES * per_leaf * Ks / (float(t.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t.item()) / Ks, "steps": Ks,
                         "scaling": "strong"}
        # ---- configs[4]: every rank takes its share of opening3's subtrees ----
        sec["c4"] = secondary_c4(sq, rank, world, dev, dist, torch, peak_mixed)
        sec["c4"]["mcts"] = secondary_c4_mcts(sq, rank, world, dev, dist, torch, 2 ** 30)
        if rank == 0 and world == 1:
            sec["c3"] = secondary_c3(sq, peak_mixed)
            sec["negascout"] = secondary_negascout(sq)
            sec.update(secondary_c0_c2(sq, int(args.c2_iterations)))
        line["secondary"] = sec

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        rate, used, secs = cpu_port_rate(boards, to_move, 12.0, threads)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d positions x %d playouts (%.1f s) of the same workload" % (N_LEAVES, used, secs),
            "note": "oracle/squava_oracle.c: C port of the Go reference's table-walk playout; no Go toolchain here"}
    if world > 1:
        dist.barrier()
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    sq.shutdown()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
