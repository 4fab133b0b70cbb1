#!/bin/bash
# round-2 GPU session Y: NegaScout child evaluation without popcounts: parity, timing split / one warp per position
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_negascout.py tests/test_gpu_host_programs.py -x -q -m gpu 2>&1 | tail -5
O=gpurun_out/r02y_negascout.txt; : > $O
for v in "SQ_NS_TRACE=1" "SQ_NS_SPLIT=0" "SQ_NS_TRACE=1"; do
  echo "== $v" >> $O; env $v SQ_NS_TRACE=1 timeout 300 python tools/bench_negascout.py 2>&1 | grep -E "^ns |positions|position," | cut -c1-230 >> $O
done
python tools/ns_sweep_table.py $O
