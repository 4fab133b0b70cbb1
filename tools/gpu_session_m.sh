#!/bin/bash
# round-2 GPU session M: killer-move ordering in the alpha-beta search: correctness, then timings with and without
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_alphabeta.py tests/test_gpu_configs.py -x -q > gpurun_out/r02m_pytest_ab.log 2>&1; echo "ab pytest rc=$?"
tail -4 gpurun_out/r02m_pytest_ab.log
OUT=gpurun_out/r02m_ab.txt; : > $OUT
echo "== killers on" >> $OUT;  AB_DEEP=1 timeout 300 python tools/ab_time.py >> $OUT 2>&1
echo "== killers off" >> $OUT; SQ_AB_NO_KILLERS=1 timeout 300 python tools/ab_time.py >> $OUT 2>&1
for dm in 12 16 20 24; do echo "== defer $dm" >> $OUT; SQ_AB_DEFER=$dm timeout 300 python tools/ab_time.py >> $OUT 2>&1; done
cat $OUT
