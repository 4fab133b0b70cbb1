#!/bin/bash
# round-2 GPU session AE: the MCTS tree in HBM (sq_mcts_decide): tests, then squavam -gpu-tree at 1e8 (and 1e9 if that is quick)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tree.py tests/test_gpu_host_programs.py -x -q -m gpu -k "tree" 2>&1 | tail -25
T=gpurun_out/r02ae_gpu_tree.txt; : > $T
for it in 100000000; do
  for w in 0 2048 32768; do
    echo "-- gpu tree $it walkers $w" >> $T
    timeout 120 squava_b200/host/squavam -C -gpu-tree -stats -i $it -walkers $w -seed 5 < /dev/null 2>&1 | grep -E "gpu tree|My move|sq_mcts" >> $T
  done
done
cat $T
if grep -q "device [0-9]\.[0-9]*s" $T; then
  echo "-- gpu tree 1e9" >> $T
  timeout 200 squava_b200/host/squavam -C -gpu-tree -stats -i 1000000000 -seed 5 < /dev/null 2>&1 | grep -E "gpu tree|My move|sq_mcts" >> $T
  tail -3 $T
fi
