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

--impl reference times that same CPU port as the reference arm (the Go reference cannot
be built here: no Go toolchain).
"""
import argparse
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
    achieved = alg_ops / (kernel_ms * 1e-3)
    traffic, warp_inst = None, None
    tp = os.path.join(ROOT, "profiles", "playout_kernel_traffic.json")
    if os.path.exists(tp) and per_leaf == PER_LEAF:
        try:
            prof = json.load(open(tp))
            traffic = prof.get("dram_bytes_per_launch")
            warp_inst = prof.get("warp_instructions_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "int32_issue", "achieved": achieved / 1e12, "peak": peak_mixed / 1e12, "unit": "Tops/s",
        "frac": achieved / peak_mixed, "traffic": traffic,
        "kernel": "sq::playout_kernel", "kernel_ms": kernel_ms, "kernel_ms_last_internal_events": last_kernel_ms,
        "algorithmic_ops_per_launch": alg_ops, "ops_model": "64 per ply + 16 per playout (SURVEY.md 8d)",
        "plies_per_launch": plies_per_step, "peak_source": "sq_int32_peak measured in this run (LOP3+IMAD interleaved)",
        "peak_alu_only_tops": peak_alu / 1e12, "peak_fma_only_tops": peak_fma / 1e12,
        "algorithmic_bytes_per_launch": N_LEAVES * (9 + 32),
    }
    if warp_inst:
        # what the kernel really issues (ncu smsp__inst_executed of the same launch) against the same peak
        roofline["executed_thread_inst_per_launch"] = warp_inst * 32
        roofline["executed_thread_inst_per_ply"] = warp_inst * 32 / plies_per_step
        roofline["frac_executed"] = warp_inst * 32 / (kernel_ms * 1e-3) / peak_mixed

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
