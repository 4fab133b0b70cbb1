#!/usr/bin/env python3
"""Run under torchrun on N GPUs: sharded playouts (NCCL merge) must equal the one-GPU call bit
for bit; root-parallel MCTS statistics merge; alpha-beta positions sharded and gathered.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/multi_gpu_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import squava_b200 as sq  # noqa: E402
from squava_b200 import dist as sqd, mcts, positions  # noqa: E402


def main():
    rank, ws, lr = sqd.world()
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    sq.init(devices=[lr], seed=positions.SEED_C2)
    boards, tm = positions.midgame_batch(1000, seed=17)
    merged = sqd.playouts_sharded(sq, boards, tm, 5000, stream_id=3, device=dev).cpu().numpy().astype(np.uint64)
    ok_playouts = None
    if rank == 0:
        single = sq.playouts(boards, tm, 5000, 3)
        ok_playouts = bool((single == merged).all())
    # alpha-beta: positions sharded by rank, value rows merged by zero-padded all-reduce
    ab_boards, _ = positions.midgame_batch(64, marks=(12, 14), seed=23)
    sc = np.array([3, 3, 0, 3, 3, 3, 4, 1, 4, 3, 0, 1, 0, 1, 0, 3, 4, 1, 4, 3, 3, 3, 0, 3, 3], np.int8)
    lo, hi = sqd.shard(64, rank, ws)
    v, bv, bm, _ = sq.alphabeta_root(ab_boards[lo:hi], 1, 10, sc)
    row = torch.from_numpy(np.where(v == sq.NO_MOVE, 0, v).astype(np.int64)).to(dev)
    full = sqd.merge_shards(row, lo, 64).cpu().numpy()
    ok_ab = None
    if rank == 0:
        v1, *_ = sq.alphabeta_root(ab_boards, 1, 10, sc)
        ok_ab = bool((np.where(v1 == sq.NO_MOVE, 0, v1) == full).all())
    # root-parallel MCTS on the GPU
    rng = np.random.Generator(np.random.PCG64(rank))
    root = mcts.uct(0, 0, -1, 20000, 1.0, mcts.gpu_rollout(sq), rng, batch=512, stream_base=100000 * (rank + 1))
    vis, wins = mcts.root_statistics(root)
    mv, mw = sqd.merge_root_statistics(vis, wins)
    best, score = mcts.choose_from_statistics(mv, mw, 1.0)
    if rank == 0:
        print(json.dumps({"world": ws, "playouts_sharded_equals_single": ok_playouts, "alphabeta_sharded_equals_single": ok_ab,
                          "root_parallel_total_visits": float(mv.sum()), "root_parallel_best_move": int(best)}), flush=True)
        assert ok_playouts and ok_ab and mv.sum() == 20000 * ws
    sq.shutdown()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
