#!/bin/bash
# round-2 GPU session B: full gpu tests on the new kernels, playout variants, alpha-beta traces
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log)
tail -5 gpurun_out/r02b_pytest.log
tools/playout_variants.sh run 2>&1 | tee gpurun_out/r02b_variants.txt
for fh in 0 1; do
  if [ $fh = 0 ]; then export SQ_AB_NO_FAST_HORIZON=1; else unset SQ_AB_NO_FAST_HORIZON; fi
  echo "== fast horizon $fh" >> gpurun_out/r02b_ab.txt
  SQ_AB_TRACE=1 python tools/bench_c4.py --skip-oracle >> gpurun_out/r02b_ab.txt 2>&1
  SQ_AB_TRACE=1 python tools/bench_ab.py >> gpurun_out/r02b_ab.txt 2>&1
  echo "-- opening3 -d 10" >> gpurun_out/r02b_ab.txt
  (time SQ_AB_TRACE=1 squava_b200/host/opening3 -d 10 > /dev/null) >> gpurun_out/r02b_ab.txt 2>&1
  echo "-- opening2 -d 12" >> gpurun_out/r02b_ab.txt
  (time squava_b200/host/opening2 -d 12) >> gpurun_out/r02b_ab.txt 2>&1
done
tail -60 gpurun_out/r02b_ab.txt
