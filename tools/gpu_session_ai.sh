#!/bin/bash
# round-2 GPU session AI: the whole gpu test suite and smoke() on the final tree
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ai_pytest.log)
tail -8 gpurun_out/r02ai_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ai_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02ai_smoke.log
