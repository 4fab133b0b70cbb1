"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the merge.

The hot path shards by unit index with no data-path collective (leaves, positions and
subtrees are independent); the only exchange is the merge of small result vectors:
one all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests) per step / decision.
"""
import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard(n, rank, world_size):
    """Contiguous share [lo, hi) of n units for `rank` -- the split the C library itself uses
    for its in-process multi-device sharding (abi.cu: share())."""
    return n * rank // world_size, n * (rank + 1) // world_size


def merge_shards(local, lo, n_total, group=None):
    """All ranks hold `local` = results for units [lo, lo+len(local)) of n_total.  Returns the
    full [n_total, ...] tensor on every rank: zero-padded all-reduce(sum)."""
    import torch
    import torch.distributed as dist
    full = torch.zeros((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    full[lo:lo + local.shape[0]] = local
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, group=group)
    return full


def merge_root_statistics(visits, wins, group=None):
    """Root-parallel MCTS: every rank grew its own tree from the same root; the per-root-move
    {visits, wins} vectors (25 + 25 float64) are summed across ranks."""
    import torch
    import torch.distributed as dist
    t = torch.stack([torch.as_tensor(visits, dtype=torch.float64), torch.as_tensor(wins, dtype=torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = t.to(dev)
        dist.all_reduce(t, group=group)
        t = t.cpu()
    return t[0].numpy(), t[1].numpy()


def playouts_sharded(sq, boards, to_move, playouts_per_leaf, stream_id=0, device=None, group=None):
    """Leaves sharded over the ranks, each rank on its own GPU, counts merged on every rank.
    Results equal the single-GPU call bit for bit (the stream is keyed by global leaf index)."""
    import torch
    rank, ws, _ = world()
    n = len(boards)
    lo, hi = shard(n, rank, ws)
    b = np.ascontiguousarray(boards[lo:hi])
    d_b = torch.from_numpy(b.view(np.uint32).reshape(-1, 2).view(np.int32).copy()).to(device)
    d_t = torch.from_numpy(np.ascontiguousarray(to_move[lo:hi], dtype=np.int8)).to(device)
    d_o = torch.zeros((hi - lo, 4), dtype=torch.int64, device=device)
    if hi > lo:
        sq.playouts_dev(d_b.data_ptr(), d_t.data_ptr(), hi - lo, playouts_per_leaf, d_o.data_ptr(),
                        leaf_index_base=lo, stream_id=stream_id, stream=torch.cuda.current_stream())
    return merge_shards(d_o, lo, n, group=group)
