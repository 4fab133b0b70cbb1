#!/bin/bash
# round-2 GPU session I (gpurun --gpus 8): bench.py under torchrun at N=2 and 8, opening3 sharded in-process over
# 1/2/4/8 devices, root-parallel squavam over 1/8 devices.  Keep it short: charged 8x.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r02i_gpus.txt
for N in 2 8; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02i_bench_n$N.json 2> gpurun_out/r02i_bench_n$N.err
  echo "bench N=$N rc=$?"; tail -c 300 gpurun_out/r02i_bench_n$N.err
done
OUT=gpurun_out/r02i_inprocess.txt; : > $OUT
for N in 1 2 4 8; do
  echo "== opening3 -d 12 -gpus $N" >> $OUT
  (time timeout 200 squava_b200/host/opening3 -d 12 -gpus $N | awk -F'\t' 'NR==1{print "elapsed printed by the program:", $4} {s+=$2} END{print "sum of values", s}') >> $OUT 2>&1
done
for N in 1 8; do
  echo "== squavam -fast -root-parallel -gpus $N" >> $OUT
  timeout 200 squava_b200/host/squavam -C -fast -root-parallel -stats -gpus $N -i 2000000000 -batch 1024 -ppl 65536 -seed 3 < /dev/null 2>&1 | grep -E "My move|fast tree|searcher" >> $OUT
done
cat $OUT
timeout 300 python -m pytest tests/test_gpu_configs.py -q -k "two_devices or overlap" > gpurun_out/r02i_pytest_multi.log 2>&1; tail -3 gpurun_out/r02i_pytest_multi.log
