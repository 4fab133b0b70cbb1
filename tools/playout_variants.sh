#!/bin/bash
# Build tuning variants of the playout kernel (macros SQ_PL_*, SQ_PLAYOUT_CTAS_PER_SM) into variants/ (git-ignored),
# here on the CPU box;  `tools/playout_variants.sh run` then times each on the GPU box with bench.py and checks
# the random stream against the emulator (the knobs must not change any result).
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
declare -A V
V[base]=""
V[shf1]="-DSQ_PL_SHF=1"
V[shf2]="-DSQ_PL_SHF=2"
V[shf4]="-DSQ_PL_SHF=4"
V[acc_alu]="-DSQ_PL_ACC_ALU=1"
V[acc_alu_shf2]="-DSQ_PL_ACC_ALU=1 -DSQ_PL_SHF=2"
V[ctas6]="-DSQ_PLAYOUT_CTAS_PER_SM=6"
if [ "$1" = "run" ]; then
    for k in "${!V[@]}"; do
        export SQUAVA_B200_LIB="$PWD/variants/libsquava_b200_$k.so"
        [ -f "$SQUAVA_B200_LIB" ] || continue
        python -m pytest tests/test_gpu_playouts.py -q -x -k "bit_exact or philox" > gpurun_out/variant_$k.test 2>&1 && ok=ok || ok=FAIL
        python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/variant_$k.json 2> gpurun_out/variant_$k.err || true
        python - "$k" "$ok" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/variant_%s.json" % sys.argv[1]))
    print("%-14s %s  %.4e playouts/s  kernel %.3f ms" % (sys.argv[1], sys.argv[2], d["value"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no result", e)
PY
    done
    exit 0
fi
for k in "${!V[@]}"; do
    nvcc ${V[$k]} -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
        -o variants/libsquava_b200_$k.so squava_b200/csrc/abi.cu 2> variants/build_$k.log &
done
wait
ls -la variants/*.so
