#!/bin/bash
# round-2 GPU session AG: the tree in HBM after the buffer-table fix: tests, then squavam -gpu-tree timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tree.py tests/test_gpu_host_programs.py -x -q -m gpu -k "tree" 2>&1 | tail -25
T=gpurun_out/r02ag_gpu_tree.txt; : > $T
for lv in 0 4; do
  echo "-- gpu tree 1e8 sync-levels $lv" >> $T
  timeout 120 squava_b200/host/squavam -C -gpu-tree -stats -i 100000000 -sync-levels $lv -seed 5 < /dev/null 2>&1 | grep -E "gpu tree|My move|sq_mcts" >> $T
done
echo "-- gpu tree 1e9 (defaults)" >> $T
timeout 300 squava_b200/host/squavam -C -gpu-tree -stats -i 1000000000 -seed 5 < /dev/null 2>&1 | grep -E "gpu tree|My move|sq_mcts" >> $T
cat $T
