#!/bin/bash
# round-2 GPU session K: full gpu tests (checked build included), configs[2] with and without the double-buffered batches
# (1e8, then the full 1e9), final ncu capture of the playout kernel for profiles/playout_kernel_traffic.json
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log)
tail -5 gpurun_out/r02k_pytest.log
T=gpurun_out/r02k_tree.txt; : > $T
echo "-- 1e8 pipelined" >> $T
timeout 300 squava_b200/host/squavam -C -fast -stats -i 100000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T
echo "-- 1e8 not pipelined" >> $T
SQUAVA_FAST_NO_PIPELINE=1 timeout 300 squava_b200/host/squavam -C -fast -stats -i 100000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T
echo "-- 1e9 pipelined" >> $T
timeout 600 squava_b200/host/squavam -C -fast -stats -i 1000000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T
cat $T
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline"
$B > gpurun_out/r02k_benchline.json 2> gpurun_out/plain2.log && ncu --set full --clock-control none --import-source on -k regex:playout_kernel -s 4 -c 1 -o gpurun_out/r02k_playout $B > gpurun_out/ncu2.log 2>&1
echo "playout full rc=$?"
