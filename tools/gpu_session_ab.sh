#!/bin/bash
# round-2 GPU session AB: bench.py at N=2 under torchrun with the final tree, then the N=1 bench line for the record
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02ab_bench_n2.json 2> gpurun_out/r02ab_bench_n2.err; echo "bench n2 rc=$?"; tail -c 400 gpurun_out/r02ab_bench_n2.err
timeout 600 python bench.py > gpurun_out/r02ab_bench_n1.json 2> gpurun_out/r02ab_bench_n1.err; echo "bench n1 rc=$?"
