#!/bin/bash
# round-2 GPU session G3: queue-driven alpha-beta, counters on their own cache lines
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02g3_ab.txt
: > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" timeout 300 python tools/ab_time.py >> $OUT 2>&1; }
run "queue mb 96 trace" SQ_AB_MIN_BUDGET=96 SQ_AB_TRACE=1
run "queue mb 384" SQ_AB_MIN_BUDGET=384 SQ_AB_TRACE=1
run "queue mb 1536" SQ_AB_MIN_BUDGET=1536 SQ_AB_TRACE=1
run "queue mb 384 noybw" SQ_AB_MIN_BUDGET=384 SQ_AB_NO_YBW=1 SQ_AB_TRACE=1
cat $OUT
