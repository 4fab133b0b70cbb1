#!/bin/bash
# round-2 GPU session J: playout variants round 3, then compute-sanitizer memcheck on the small mixed case
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
tools/playout_variants.sh run 2>&1 | tee gpurun_out/r02j_variants.txt
timeout 120 python tools/sanitize_case.py > gpurun_out/r02j_sanitize_plain.log 2>&1; echo "plain rc=$?"
timeout 900 compute-sanitizer --tool ${1:-memcheck} --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r02j_sanitize_${1:-memcheck}.log 2>&1; echo "sanitizer ${1:-memcheck} rc=$?"
tail -15 gpurun_out/r02j_sanitize_${1:-memcheck}.log
