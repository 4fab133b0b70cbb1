#!/bin/bash
# round-2 GPU session T: NegaScout split, items in asynchronous flights: parity, then a sweep of the knobs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_negascout.py -x -q -m gpu 2>&1 | tail -15
O=gpurun_out/r02t_negascout_sweep.txt; : > $O
for v in "SQ_NS_TRACE=1" "SQ_NS_FLIGHTS=1" "SQ_NS_FLIGHTS=4" "SQ_NS_FLIGHTS=16" "SQ_NS_ITEM_BUDGET=4096" "SQ_NS_ITEM_BUDGET=1024" \
         "SQ_NS_ITEM_BUDGET=4096 SQ_NS_FLIGHTS=16" "SQ_NS_ITEM_BUDGET=4096 SQ_NS_SPLIT_PLY=2 SQ_NS_SPLIT_PLY_NULL=1" \
         "SQ_NS_ITEM_BUDGET=4096 SQ_NS_BUDGET=32768" "SQ_NS_ITEM_BUDGET=4096 SQ_NS_DIRECT=100000"; do
  echo "== $v" >> $O; env $v SQ_NS_TRACE=1 timeout 300 python tools/bench_negascout.py 2>&1 | grep -E "^ns split|positions|position," | cut -c1-230 >> $O
done
cat $O
