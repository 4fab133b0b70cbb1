#!/bin/bash
# round-2 GPU session AD: bench.py at N=4 under torchrun with the final tree
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02ad_bench_n4.json 2> gpurun_out/r02ad_bench_n4.err; echo "bench n4 rc=$?"; tail -c 300 gpurun_out/r02ad_bench_n4.err
