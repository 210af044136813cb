#!/bin/bash
# Lean A/B of the aggregation kernel's (resident CTAs, pipeline depth) variants (FIELDCONV_B200_AGG_VARIANT=b0,b1,b2).
#   gpurun --timeout 420 -- 'bash tools/gpu_variants.sh r01h'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run() {   # variant-string  layer_bench args...
  local v=$1; shift
  FIELDCONV_B200_AGG_VARIANT=$v timeout 100 python tools/layer_bench.py "$@" --tag "var$v"
}
{
  for v in 32,32,32 32,32,31 32,32,41 32,32,22; do run $v --side 284 --channels 48 --band 2 --rings 6; done
  for v in 32,32,22 32,32,31 32,32,32; do run $v --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p; done
  for v in 32,32,32 32,31,32 32,42,32 32,41,32; do run $v --side 1000 --channels 32 --band 1 --rings 6 --steps 5; done
  for v in 32,32,32 32,31,32 32,41,32; do run $v --side 284 --channels 128 --band 1 --rings 6; done
  for v in 32,32,32 32,41,32; do run $v --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --precision 2xf16; done
} > $OUT/${TAG}_variants.jsonl 2> $OUT/${TAG}_variants.err
python - <<PY
import json
for l in open("$OUT/${TAG}_variants.jsonl"):
    d = json.loads(l)
    k = d["kernels_ms"]
    print(d["tag"], d["vertices"], d["channels"], d["band_limit"], d["precision"], "ms", d["ms_fwd_bwd"],
          {n: v for n, v in k.items() if n.startswith("aggregate")})
PY
tail -3 $OUT/${TAG}_variants.err
