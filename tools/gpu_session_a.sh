#!/bin/bash
# round-2 GPU session A: tests, bench with the secondary block, ncu launch list + full captures
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.csv
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log)
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline"
$B > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02a_launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:playout_kernel -s 4 -c 1 -o gpurun_out/r02a_playout $B > gpurun_out/ncu2.log 2>&1
echo "playout full rc=$?"
A="python tools/bench_c4.py --skip-oracle"
$A > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ab_dfs -c 14 -o gpurun_out/r02a_abdfs $A > gpurun_out/ncu3.log 2>&1
echo "abdfs full rc=$?"
ls -la gpurun_out
