#!/bin/bash
# round-2 GPU session AK: ncu launch list of the tree in HBM (per-kernel durations of a few batches)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
A="squava_b200/host/squavam -C -gpu-tree -stats -i 3000000 -seed 4"
timeout 60 $A < /dev/null > gpurun_out/plain_tree.log 2>&1 && timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ak_tree_launches.csv $A < /dev/null > gpurun_out/ncu_tree.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_tree.log
