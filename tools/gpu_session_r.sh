#!/bin/bash
# round-2 GPU session R: NegaScout split over many warps (NsSplit): parity, then timing against the one-warp replay
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_negascout.py -x -q -m gpu 2>&1 | tail -15
echo "== split (default)"; SQ_NS_TRACE=1 timeout 600 python tools/bench_negascout.py > gpurun_out/r02r_negascout_split.json 2> gpurun_out/r02r_split.err; tail -12 gpurun_out/r02r_split.err
echo "== one warp per position"; SQ_NS_SPLIT=0 timeout 600 python tools/bench_negascout.py > gpurun_out/r02r_negascout_plain.json 2> gpurun_out/r02r_plain.err; tail -3 gpurun_out/r02r_plain.err
cat gpurun_out/r02r_negascout_split.json; echo; cat gpurun_out/r02r_negascout_plain.json
for v in "SQ_NS_SPLIT_PLY=2" "SQ_NS_SPLIT_PLY=4" "SQ_NS_SPLIT_PLY=4 SQ_NS_SPLIT_PLY_NULL=3" "SQ_NS_SPLIT_PLY=5 SQ_NS_SPLIT_PLY_NULL=2" "SQ_NS_BUDGET=32768" "SQ_NS_BUDGET=524288"; do
  echo "== $v"; env $v SQ_NS_TRACE=1 timeout 600 python tools/bench_negascout.py 2>&1 | grep -E "^ns split|positions|position," | cut -c1-200
done
