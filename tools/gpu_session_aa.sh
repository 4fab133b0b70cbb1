#!/bin/bash
# round-2 GPU session AA: NegaScout items land one by one (landing lists read while the kernels run): full gpu tests, timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aa_pytest.log)
tail -5 gpurun_out/r02aa_pytest.log
O=gpurun_out/r02aa_negascout.txt; : > $O
for v in "SQ_NS_TRACE=1" "SQ_NS_FLIGHTS=16" "SQ_NS_ITEM_BUDGET=4096" "SQ_NS_ITEM_BUDGET=16384"; do
  echo "== $v" >> $O; env $v SQ_NS_TRACE=1 timeout 300 python tools/bench_negascout.py 2>&1 | grep -E "^ns |positions|position," | cut -c1-230 >> $O
done
python tools/ns_sweep_table.py $O
