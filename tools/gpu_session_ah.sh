#!/bin/bash
# round-2 GPU session AH: the tree tests (all of them, no -x), the host-program test, bench.py for the record
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tree.py tests/test_gpu_host_programs.py -q -m gpu -k "tree" 2>&1 | tail -30
timeout 600 python bench.py > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02ah_bench.err
