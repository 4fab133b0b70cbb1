#!/bin/bash
# round-2 GPU session N2: one vs two killers per depth; brothers-wait passes on top of the killers
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02n2_ab.txt; : > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" AB_DEEP=1 timeout 300 python tools/ab_time.py >> $OUT 2>&1; }
run "two killers (default)" X=1
run "one killer" SQ_AB_ONE_KILLER=1
run "two killers, brothers wait" SQ_AB_YBW=1
cat $OUT
timeout 600 python -m pytest tests/test_gpu_alphabeta.py -x -q 2>&1 | tail -3
