#!/bin/bash
# round-2 GPU session G: queue-driven alpha-beta with polling back-off; in-process timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r02g_ab.txt
: > $OUT
run() { echo "== $1" >> $OUT; shift; env "$@" timeout 300 python tools/ab_time.py >> $OUT 2>&1; }
run "passes default" SQ_AB_PASSES=1
run "passes budget 64" SQ_AB_BUDGET=64
run "queue mb 96" SQ_AB_MIN_BUDGET=96
run "queue mb 384" SQ_AB_MIN_BUDGET=384
run "queue mb 1536" SQ_AB_MIN_BUDGET=1536
run "queue mb 384 noybw" SQ_AB_MIN_BUDGET=384 SQ_AB_NO_YBW=1
run "queue mb 1536 noybw" SQ_AB_MIN_BUDGET=1536 SQ_AB_NO_YBW=1
cat $OUT
