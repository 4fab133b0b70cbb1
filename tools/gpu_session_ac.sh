#!/bin/bash
# round-2 GPU session AC: ncu capture of the final NegaScout kernels (no popcounts, landing lists), NegaScout tests once more
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_negascout.py tests/test_gpu_checked_build.py -x -q -m gpu 2>&1 | tail -3
A="python tools/ncu_negascout.py"
$A > gpurun_out/plain_ns2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:negascout_fast -c 12 -o gpurun_out/r02ac_negascout $A > gpurun_out/ncu_ns2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_ns2.log
