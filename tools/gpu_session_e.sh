#!/bin/bash
# round-2 GPU session E: queue-driven alpha-beta after the L2-read / throttling fixes
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02e_ab.txt
echo "== correctness" > $OUT
timeout 900 python -m pytest tests/test_gpu_alphabeta.py tests/test_gpu_configs.py -x -q > gpurun_out/r02e_pytest_ab.log 2>&1; echo "ab pytest rc=$?" | tee -a $OUT
tail -5 gpurun_out/r02e_pytest_ab.log
for mode in queue queue_noybw; do
  unset SQ_AB_NO_YBW SQ_AB_PASSES
  [ $mode = queue_noybw ] && export SQ_AB_NO_YBW=1
  echo "== $mode" >> $OUT
  SQ_AB_TRACE=1 timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  echo "-- opening3 -d 10" >> $OUT
  (time SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening3 -d 10 2>&1 >/dev/null ) >> $OUT 2>&1
  echo "-- opening3 -d 12" >> $OUT
  (time SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening3 -d 12 2>&1 >/dev/null ) >> $OUT 2>&1
  echo "-- opening2 -d 12" >> $OUT
  (time SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening2 -d 12 2>&1 ) >> $OUT 2>&1
done
unset SQ_AB_NO_YBW SQ_AB_PASSES
for mb in 24 256 1024; do
  echo "== queue min_budget $mb" >> $OUT
  SQ_AB_MIN_BUDGET=$mb timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  (time SQ_AB_MIN_BUDGET=$mb SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening2 -d 12 2>&1 | grep "queue:") >> $OUT 2>&1
  (time SQ_AB_MIN_BUDGET=$mb SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening3 -d 10 2>&1 | grep "queue:") >> $OUT 2>&1
done
cat $OUT
echo "== host tree knobs" > gpurun_out/r02e_tree.txt
for cfg in "8 4" "4 2" "2 1" "4 8"; do
  set -- $cfg
  echo "-- group $1 ahead $2" >> gpurun_out/r02e_tree.txt
  SQUAVA_FAST_GROUP=$1 SQUAVA_FAST_AHEAD=$2 timeout 300 squava_b200/host/squavam -C -fast -stats -i 100000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> gpurun_out/r02e_tree.txt
done
cat gpurun_out/r02e_tree.txt
