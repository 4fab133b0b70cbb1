#!/usr/bin/env python3
"""configs[4]: "opening3 deep evaluation of 6 unique first moves x all unique replies,
root-parallel MCTS + alpha-beta sharded across 2/4/8 GPUs with NCCL merge".

Run under torchrun (or plain python for 1 GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/bench_c5.py [--depth 10] [--check-depth 7]

Part A  alpha-beta: opening3's 85 subtrees (dialect AB-3) sharded round-robin over the ranks,
        value vector merged with one all-reduce; --check-depth also runs a shallower depth and
        compares every value with the CPU oracle.
Part B  root-parallel MCTS from the same 85 two-mark positions: every rank grows its own trees
        (different host seeds, disjoint Philox streams), per-root-move {visits, wins} vectors
        are summed with one all-reduce per position batch.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import squava_b200 as sq  # noqa: E402
from squava_b200 import dist as sqd, mcts  # noqa: E402

SCORES_OPENING = np.array([3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 2, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3], np.int8)
FIRST_MOVES = [0, 1, 2, 6, 7, 12]


def d4_images(c):
    x, y = divmod(c, 5)
    return [5 * a + b for a, b in ((x, y), (y, 4 - x), (4 - x, 4 - y), (4 - y, x), (4 - x, y), (x, 4 - y), (4 - y, 4 - x), (y, x))]


def opening3_pairs():
    """opening3.go:66-128: (first, reply) pairs not equivalent under D4 to an earlier one."""
    done, out = set(), []
    for f in FIRST_MOVES:
        for r in range(25):
            if r == f:
                continue
            if any((a, b) in done for a, b in zip(d4_images(f), d4_images(r))):
                continue
            done.add((f, r))
            out.append((f, r))
    return out


def subtrees(pairs):
    t = np.zeros(len(pairs), sq.SUBTREE)
    for i, (f, r) in enumerate(pairs):
        t[i] = (1 << f, 1 << r, r, 2, 1, 0, int(SCORES_OPENING[f]))
    return t


def ab_sharded(pairs, depth, rank, ws, dev):
    mine = list(range(rank, len(pairs), ws))
    vals = torch.zeros(len(pairs), dtype=torch.int64, device=dev)
    nodes = torch.zeros(len(pairs), dtype=torch.int64, device=dev)
    t0 = time.perf_counter()
    if mine:
        v, n = sq.alphabeta_subtrees(subtrees([pairs[i] for i in mine]), sq.AB_OPENING3, depth, SCORES_OPENING)
        vals[mine] = torch.from_numpy(v.astype(np.int64)).to(dev)
        nodes[mine] = torch.from_numpy(n.astype(np.int64)).to(dev)
    if ws > 1:
        dist.all_reduce(vals)
        dist.all_reduce(nodes)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return vals.cpu().numpy(), int(nodes.sum().item()), float(dt.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--check-depth", type=int, default=7)
    ap.add_argument("--mcts-iterations", type=int, default=2_000_000, help="playouts per position per rank")
    args = ap.parse_args()
    rank, ws, lr = sqd.world()
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    sq.init(devices=[lr], seed=0xC5)
    pairs = opening3_pairs()
    assert len(pairs) == 85
    rec = {"config": "configs[4]", "world": ws, "subtrees": len(pairs)}
    # ---- part A ---------------------------------------------------------------------
    if args.check_depth:
        v, _, _ = ab_sharded(pairs, args.check_depth, rank, ws, dev)
        if rank == 0:
            import oracle_lib
            o = oracle_lib.load()
            sc = o.tables()["scores_opening"]
            assert o.opening3_subtrees() == pairs
            want = [o.ab_subtree(3, 1 << f, 1 << r, r, 2, 1, sc[f], args.check_depth, sc)[0] for f, r in pairs]
            rec["alphabeta_check_depth"] = args.check_depth
            rec["alphabeta_values_match_oracle"] = bool(v.tolist() == want)
    v, nodes, secs = ab_sharded(pairs, args.depth, rank, ws, dev)
    rec.update({"alphabeta_depth": args.depth, "alphabeta_seconds": secs, "alphabeta_gpu_evaluator_calls": nodes,
                "alphabeta_values": v.tolist()})
    # ---- part B ---------------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(1000 + rank))
    roll = mcts.gpu_rollout(sq)
    t0 = time.perf_counter()
    best = []
    for i, (f, r) in enumerate(pairs):
        root = mcts.uct(1 << f, 1 << r, -1, args.mcts_iterations, 1.0, roll, rng, batch=256, ppl=1024,
                        stream_base=(rank * 85 + i) << 20)
        vis, wins = mcts.root_statistics(root)
        mv, mw = sqd.merge_root_statistics(vis, wins)
        best.append(int(mcts.choose_from_statistics(mv, mw, 1.0)[0]))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    rec.update({"mcts_playouts_per_position": args.mcts_iterations * ws, "mcts_seconds": float(dt.item()),
                "mcts_playouts_per_s": 85 * args.mcts_iterations * ws / float(dt.item()), "mcts_best_moves": best})
    if rank == 0:
        print(json.dumps(rec), flush=True)
    sq.shutdown()
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
