"""world_size-2 gloo tests of the multi-GPU host logic (no device): shard arithmetic, the
zero-padded all-reduce merge, and root-parallel MCTS statistics merging.  The rollout is a
CPU stand-in (the Python emulation of the kernel) -- test scaffolding only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import playout_emulator as E


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def emu_rollout(boards, to_move, ppl, stream_id):
    out = []
    for i, ((x, o), tm) in enumerate(zip(boards, to_move)):
        terminal = E.owns(x, E.TRIPLETS) or E.owns(o, E.TRIPLETS) or (x | o) == 0x1ffffff
        if terminal:
            w = 1 if E.owns(x, E.QUADS) else -1 if E.owns(o, E.QUADS) else -1 if E.owns(x, E.TRIPLETS) else 1 if E.owns(o, E.TRIPLETS) else 0
            out.append([ppl if w == 1 else 0, ppl if w == -1 else 0, ppl if w == 0 else 0, 0])
        else:
            out.append(E.emulate_item(x, o, tm, ppl, i, 0, 99, stream_id))
    return out


def worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from squava_b200 import dist as sqd, mcts
    # 1. shard + merge
    n = 37
    lo, hi = sqd.shard(n, rank, ws)
    local = torch.arange(lo, hi, dtype=torch.int64).unsqueeze(1).repeat(1, 4) + 1
    full = sqd.merge_shards(local, lo, n)
    ok_merge = bool((full[:, 0] == torch.arange(1, n + 1)).all())
    # 2. root-parallel MCTS: different seeds per rank, merged root statistics identical on all ranks
    rng = np.random.Generator(np.random.PCG64(100 + rank))
    x, o = 0x0412825, 0x098c0c0
    root = mcts.uct(x, o, -1, 600, 1.0, emu_rollout, rng, batch=16, stream_base=1000 * rank)
    v, w = mcts.root_statistics(root)
    mv, mw = sqd.merge_root_statistics(v, w)
    best, _ = mcts.choose_from_statistics(mv, mw, 1.0)
    # 3. size-sharded subtrees (bench.py secondary_c4): every rank measures DIFFERENT pilot weights (evaluator counts
    #    depend on GPU scheduling); after the broadcast of rank 0's they must derive the same partition, and the
    #    zero-padded all-reduce must then hold every subtree's value exactly once
    import bench
    local_w = torch.tensor(np.random.Generator(np.random.PCG64(7 + rank)).integers(1, 1000, size=85), dtype=torch.int64)
    dist.broadcast(local_w, src=0)
    mine = bench.lpt_assign([float(x) for x in local_w.numpy()], ws)[rank]
    vals = torch.zeros(85, dtype=torch.int64)
    vals[mine] = torch.tensor([1000 + i for i in mine], dtype=torch.int64)
    dist.all_reduce(vals)
    ok_merge = ok_merge and vals.tolist() == [1000 + i for i in range(85)]
    q.put((rank, ok_merge, float(v.sum()), float(mv.sum()), mv.tolist(), best))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_and_root_parallel_mcts():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, ok0, own0, tot0, mv0, best0), (r1, ok1, own1, tot1, mv1, best1) = res
    assert ok0 and ok1
    assert own0 == 600 and own1 == 600 and tot0 == tot1 == 1200     # visits add up across trees
    assert mv0 == mv1 and best0 == best1                              # every rank sees the same merged vector
    assert (0x0412825 | 0x098c0c0) >> best0 & 1 == 0                  # and picks a legal move


def test_shard_covers_everything():
    from squava_b200 import dist as sqd
    for n in (0, 1, 7, 4096, 4097):
        for ws in (1, 2, 3, 8):
            spans = [sqd.shard(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_python_uct_matches_reference_shape():
    """sequential mode (batch=1): every iteration adds exactly one visit to the root, children are
    created one per iteration until the root is fully expanded (mcts.go:158-171)."""
    from squava_b200 import mcts
    rng = np.random.Generator(np.random.PCG64(1))
    root = mcts.uct(0, 0, -1, 40, 1.0, emu_rollout, rng, batch=1)
    assert root.visits == 40 and len(root.children) == 25
    assert sum(c.visits for c in root.children) == 40
    assert all(0.0 <= c.wins <= c.visits for c in root.children)
