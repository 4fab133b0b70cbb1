#!/bin/bash
# round-2 GPU session AL: the tree's decide step in single precision: tree tests, 1e8 and 1e9 timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_tree.py -q -m gpu 2>&1 | tail -4
T=gpurun_out/r02al_gpu_tree.txt; : > $T
echo "-- gpu tree 1e8 (single-precision decide)" >> $T
timeout 20 squava_b200/host/squavam -C -gpu-tree -stats -i 100000000 -seed 5 < /dev/null 2>&1 | grep -E "gpu tree|My move|sq_mcts" >> $T
echo "-- gpu tree 1e9" >> $T
timeout 45 squava_b200/host/squavam -C -gpu-tree -stats -i 1000000000 -seed 5 < /dev/null 2>&1 | grep -E "gpu tree|My move|sq_mcts" >> $T
cat $T
