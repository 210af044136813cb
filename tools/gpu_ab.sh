#!/bin/bash
# A/B of environment switches on the default bench (no cfg 4, no CPU baseline): one line per setting.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh r02r "FIELDCONV_B200_GEMM_RAGGED=0" "FIELDCONV_B200_GEMM_REVERSE=0"'
TAG=${1:-rXX}; shift
OUT=gpurun_out
mkdir -p $OUT
run() {
  env "$@" python bench.py --skip-cfg4 --skip-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'env': sys.argv[1:], 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'kernels_ms': {k: v['ms'] for k, v in d['kernel_shares'].items()}}))" "$@"
}
: > $OUT/${TAG}_ab.jsonl
run X=0 | tee -a $OUT/${TAG}_ab.jsonl
for s in "$@"; do run $s | tee -a $OUT/${TAG}_ab.jsonl; done
run X=0 | tee -a $OUT/${TAG}_ab.jsonl
