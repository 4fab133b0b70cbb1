#!/bin/bash
# round-2 GPU session C: gpu tests (YBW + fast horizon), playout variants round 2, alpha-beta traces, bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log)
tail -5 gpurun_out/r02c_pytest.log
tools/playout_variants.sh run 2>&1 | tee gpurun_out/r02c_variants.txt
for mode in ybw noybw; do
  if [ $mode = noybw ]; then export SQ_AB_NO_YBW=1; else unset SQ_AB_NO_YBW; fi
  echo "== $mode" >> gpurun_out/r02c_ab.txt
  SQ_AB_TRACE=1 python tools/bench_c4.py --skip-oracle 2>&1 | grep -v "^\[ab\] dfs" >> gpurun_out/r02c_ab.txt
  python tools/bench_c4.py --skip-oracle >> gpurun_out/r02c_ab.txt 2>&1
  echo "-- opening3 -d 10" >> gpurun_out/r02c_ab.txt
  (time SQ_AB_TRACE=1 squava_b200/host/opening3 -d 10 2>&1 >/dev/null | grep -v "^\[ab\] dfs" ) >> gpurun_out/r02c_ab.txt 2>&1
  echo "-- opening3 -d 12" >> gpurun_out/r02c_ab.txt
  (time SQ_AB_TRACE=1 squava_b200/host/opening3 -d 12 2>&1 >/dev/null | grep -v "^\[ab\] dfs" ) >> gpurun_out/r02c_ab.txt 2>&1
  echo "-- opening2 -d 12" >> gpurun_out/r02c_ab.txt
  (time squava_b200/host/opening2 -d 12) >> gpurun_out/r02c_ab.txt 2>&1
done
cat gpurun_out/r02c_ab.txt
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
