#!/bin/bash
# round-2 GPU session N: move ordering variants, budgets and deferred horizon evaluations on top of the killer moves
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02n_ab.txt; : > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" AB_DEEP=1 timeout 300 python tools/ab_time.py >> $OUT 2>&1; }
run "default (killer)" X=1
run "order 1 (killer, then blocks)" SQ_AB_ORDER=1
run "order 2 (blocks, then killer)" SQ_AB_ORDER=2
run "defer 16" SQ_AB_DEFER=16
run "defer 22" SQ_AB_DEFER=22
run "budget 128" SQ_AB_BUDGET=128
run "budget 1024" SQ_AB_BUDGET=1024
cat $OUT
