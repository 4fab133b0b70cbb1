#!/bin/bash
# round-2 GPU session U: NegaScout split, given-up items resume where they stood: parity, then a sweep of the knobs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_negascout.py -x -q -m gpu 2>&1 | tail -15
O=gpurun_out/r02u_negascout_sweep.txt; : > $O
for v in "SQ_NS_FLIGHTS=16" "SQ_NS_FLIGHTS=8" "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=8192" "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=4096" "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=2048" \
         "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=1024" "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=4096 SQ_NS_SPLIT_PLY=1" "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=4096 SQ_NS_BUDGET=32768" \
         "SQ_NS_FLIGHTS=16 SQ_NS_ITEM_BUDGET=4096 SQ_NS_BUDGET=524288"; do
  echo "== $v" >> $O; env $v SQ_NS_TRACE=1 timeout 300 python tools/bench_negascout.py 2>&1 | grep -E "^ns split|positions|position," | cut -c1-230 >> $O
done
cat $O
