#!/bin/bash
# round-2 GPU session Z: validation of the final tree -- full gpu tests, smoke, bench (N=1)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log)
tail -5 gpurun_out/r02z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02z_smoke.log
timeout 600 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02z_bench.err
