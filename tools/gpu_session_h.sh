#!/bin/bash
# round-2 GPU session H: full gpu test suite, deep alpha-beta timings, bench, ncu captures of the current kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log)
tail -5 gpurun_out/r02h_pytest.log
AB_DEEP=1 timeout 600 python tools/ab_time.py > gpurun_out/r02h_ab_deep.txt 2>&1; cat gpurun_out/r02h_ab_deep.txt
timeout 600 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02h_bench.err
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline"
$B > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02h_launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:playout_kernel -s 4 -c 1 -o gpurun_out/r02h_playout $B > gpurun_out/ncu2.log 2>&1
echo "playout full rc=$?"
A="python tools/bench_c4.py --skip-oracle"
$A > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ab_dfs -c 14 -o gpurun_out/r02h_abdfs $A > gpurun_out/ncu3.log 2>&1
echo "abdfs full rc=$?"
ls -la gpurun_out | tail -12
