#!/bin/bash
# round-2 GPU session P: arena tree A/B on the GPU box -- 24-byte nodes + gathers vs 16-byte nodes + shuffles + fast log
cd "$GRAFT_REPO_ROOT"
export LD_LIBRARY_PATH=$PWD/squava_b200:$LD_LIBRARY_PATH
mkdir -p gpurun_out
T=gpurun_out/r02p_tree.txt; : > $T
for v in old24 new16 old24 new16; do
  echo "-- $v 1e8" >> $T
  timeout 300 variants/squavam_$v -C -fast -stats -i 100000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T
done
for v in old24 new16; do
  echo "-- $v 4e8" >> $T
  timeout 400 variants/squavam_$v -C -fast -stats -i 400000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T
done
cat $T
