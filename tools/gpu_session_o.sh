#!/bin/bash
# round-2 GPU session O: final validation -- full gpu tests, smoke, bench (N=1), ncu capture of the alpha-beta passes
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log)
tail -5 gpurun_out/r02o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02o_smoke.log
timeout 600 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02o_bench.err
A="python tools/bench_c4.py --skip-oracle"
$A > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ab_dfs -c 14 -o gpurun_out/r02o_abdfs $A > gpurun_out/ncu3.log 2>&1
echo "abdfs full rc=$?"
