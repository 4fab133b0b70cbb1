#!/bin/bash
# round-2 GPU session F: queue-driven alpha-beta after the publish-order fix; pass-mode budget sweep
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02f_ab.txt
echo "== correctness" > $OUT
timeout 900 python -m pytest tests/test_gpu_alphabeta.py tests/test_gpu_configs.py -x -q > gpurun_out/r02f_pytest_ab.log 2>&1; echo "ab pytest rc=$?" | tee -a $OUT
tail -5 gpurun_out/r02f_pytest_ab.log
SQ_AB_PASSES=1 squava_b200/host/opening2 -d 12 | cut -f1,2 > gpurun_out/o2_pass.txt
SQ_AB_PASSES=1 squava_b200/host/opening3 -d 10 | cut -f1,2 | grep "^<" > gpurun_out/o3_pass.txt
for mode in queue queue_noybw; do
  unset SQ_AB_NO_YBW SQ_AB_PASSES
  [ $mode = queue_noybw ] && export SQ_AB_NO_YBW=1
  echo "== $mode" >> $OUT
  SQ_AB_TRACE=1 timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  echo "-- opening3 -d 10" >> $OUT
  (time SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening3 -d 10 2>>$OUT | cut -f1,2 | grep "^<" > gpurun_out/o3_q.txt ) >> $OUT 2>&1
  diff gpurun_out/o3_pass.txt gpurun_out/o3_q.txt > /dev/null && echo "opening3 values equal" >> $OUT || echo "opening3 values DIFFER" >> $OUT
  echo "-- opening3 -d 12" >> $OUT
  (time SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening3 -d 12 2>&1 >/dev/null ) >> $OUT 2>&1
  echo "-- opening2 -d 12" >> $OUT
  (time SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening2 -d 12 2>>$OUT | cut -f1,2 > gpurun_out/o2_q.txt ) >> $OUT 2>&1
  diff gpurun_out/o2_pass.txt gpurun_out/o2_q.txt > /dev/null && echo "opening2 values equal" >> $OUT || echo "opening2 values DIFFER" >> $OUT
done
unset SQ_AB_NO_YBW SQ_AB_PASSES
for mb in 32 384 1536; do
  echo "== queue min_budget $mb" >> $OUT
  SQ_AB_MIN_BUDGET=$mb timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  (time SQ_AB_MIN_BUDGET=$mb SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening2 -d 12 2>&1 | grep "queue:") >> $OUT 2>&1
  (time SQ_AB_MIN_BUDGET=$mb SQ_AB_TRACE=1 timeout 120 squava_b200/host/opening3 -d 10 2>&1 | grep "queue:") >> $OUT 2>&1
done
for b in 64 128 512; do
  echo "== passes budget $b" >> $OUT
  SQ_AB_BUDGET=$b timeout 120 python tools/bench_c4.py --skip-oracle >> $OUT 2>&1
  (time SQ_AB_BUDGET=$b timeout 120 squava_b200/host/opening2 -d 12 > /dev/null) >> $OUT 2>&1
  (time SQ_AB_BUDGET=$b timeout 120 squava_b200/host/opening3 -d 10 > /dev/null) >> $OUT 2>&1
done
cat $OUT
