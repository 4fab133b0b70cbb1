#!/bin/bash
# round-2 GPU session L (gpurun --gpus 2): bench.py under torchrun at N=2 after the partition fix
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err
echo "bench N=2 rc=$?"; tail -c 300 gpurun_out/r02l_bench_n2.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_reference_arm.json 2>&1; echo "reference arm rc=$?"
