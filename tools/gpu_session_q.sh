#!/bin/bash
# round-2 GPU session Q: NegaScout root moves in waves: parity, then timing with and without
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_negascout.py -x -q -m gpu 2>&1 | tail -15
echo "== waves"; timeout 600 python tools/bench_negascout.py > gpurun_out/r02q_negascout_waves.json 2> gpurun_out/r02q_waves.err; tail -3 gpurun_out/r02q_waves.err
echo "== one warp per position"; SQ_NS_NO_WAVES=1 timeout 600 python tools/bench_negascout.py > gpurun_out/r02q_negascout_nowaves.json 2> gpurun_out/r02q_nowaves.err; tail -3 gpurun_out/r02q_nowaves.err
cat gpurun_out/r02q_negascout_waves.json; echo; cat gpurun_out/r02q_negascout_nowaves.json
