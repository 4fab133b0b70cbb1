#!/bin/bash
# round-2 GPU session V: validation after the NegaScout split and the 16-byte arena nodes -- full gpu tests, smoke, bench
# (N=1), configs[2] at 1e9, ncu capture of the NegaScout item kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest.log)
tail -5 gpurun_out/r02v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02v_smoke.log
timeout 600 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02v_bench.err
T=gpurun_out/r02v_tree_1e9.txt; : > $T
timeout 600 squava_b200/host/squavam -C -fast -stats -i 1000000000 -batch 65536 -seed 5 < /dev/null 2>&1 | grep -E "fast tree|My move" >> $T; cat $T
A="python tools/ncu_negascout.py"
$A > gpurun_out/plain_ns.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:negascout_fast -c 12 -o gpurun_out/r02v_negascout $A > gpurun_out/ncu_ns.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_ns.log
