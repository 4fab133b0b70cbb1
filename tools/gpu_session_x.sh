#!/bin/bash
# round-2 GPU session X: NegaScout split: the child an item gave up in restarts as a loop (SQ_NS_BIG_HINT): parity, sweep
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_negascout.py -x -q -m gpu 2>&1 | tail -15
O=gpurun_out/r02x_negascout_sweep.txt; : > $O
for v in "SQ_NS_TRACE=1" "SQ_NS_BIG_HINT=0" "SQ_NS_ITEM_BUDGET=16384" "SQ_NS_ITEM_BUDGET=4096" "SQ_NS_ITEM_BUDGET=16384 SQ_NS_ITEM_BUDGET_WIDE=2048" "SQ_NS_BUDGET=32768"; do
  echo "== $v" >> $O; env $v SQ_NS_TRACE=1 timeout 300 python tools/bench_negascout.py 2>&1 | grep -E "^ns |positions|position," | cut -c1-230 >> $O
done
python tools/ns_sweep_table.py $O
