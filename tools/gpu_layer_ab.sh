#!/bin/bash
# A/B of environment switches on single-layer points (cfg-2 layer and 1 M x C=32): per setting the kernel times.
#   gpurun --timeout 600 -- 'bash tools/gpu_layer_ab.sh r04b "FIELDCONV_B200_AGG_BLOCK=128" "FIELDCONV_B200_AGG_BLOCK=192"'
TAG=${1:-rXX}; shift
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_layer_ab.jsonl
run() {
  for cfg in "--side 284 --channels 48 --band 2 --rings 6" "--side 1000 --channels 32 --band 1 --rings 6 --steps 5"; do
    env "$@" python tools/layer_bench.py $cfg 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels_ms') or d.get('kernel_ms') or {}
print(json.dumps({'env': sys.argv[1:], 'n': d.get('vertices'), 'c': d.get('channels'), 'ms': d.get('ms_fwd_bwd') or d.get('ms'), 'kernels': k}))" "$@" | tee -a $OUT/${TAG}_layer_ab.jsonl
  done
}
run X=0
for s in "$@"; do run $s; done
run X=0
