#!/bin/bash
# round-2 GPU session P2: arena tree with 16-byte nodes: UCB scan by loads + shuffles / gathers / scalar
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=gpurun_out/r02p2_tree.txt; : > $T
for v in shuffle gather scalar shuffle gather; do
  echo "-- scan $v 1e8" >> $T
  SQUAVA_FAST_SCAN=$v timeout 300 squava_b200/host/squavam -C -fast -stats -i 100000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T
done
cat $T
