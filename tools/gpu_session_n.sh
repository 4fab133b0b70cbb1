#!/bin/bash
# round-2 GPU session N3: brothers wait only under the top k plies of the split tree
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02n3_ab.txt; : > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" AB_DEEP=1 timeout 400 python tools/ab_time.py >> $OUT 2>&1; }
run "ybw top 1" SQ_AB_YBW=1
run "ybw top 2" SQ_AB_YBW=2
run "ybw top 3" SQ_AB_YBW=3
run "ybw top 5" SQ_AB_YBW=5
cat $OUT
