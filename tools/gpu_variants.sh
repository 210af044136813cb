#!/bin/bash
# Lean A/B of the aggregation kernel's variants (FIELDCONV_B200_AGG_VARIANT=b0,b1,b2; code = 10*CTAs/SM + pipeline depth,
# +100 = FAST arithmetic).  Parity of a variant: FIELDCONV_B200_AGG_VARIANT=... python tools/variant_probe.py
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
  # band limit 2, fp32 output (the cfg-2 bench path): default 31 vs depth 3 / FAST arithmetic
  for v in 32,41,31 32,41,33 32,41,34 32,41,131 32,41,133 32,41,134 32,41,43 32,41,44; do run $v --side 284 --channels 48 --band 2 --rings 6; done
  # band limit 2, packed output: default 22 vs out-of-line ring store at 3 CTAs/SM (packed wins at B=2 if this gets close to the fp32 kernel)
  for v in 32,32,22 32,32,231 32,32,232 32,32,331; do run $v --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p; done
  # band limit 1, packed output (1 M vertices, C=32): default 32 vs depth 3 / FAST
  for v in 32,32,31 32,33,31 32,34,31 32,44,31 32,132,31 32,134,31; do run $v --side 1000 --channels 32 --band 1 --rings 6 --steps 5; done
  # band limit 1, fp32 output: default 41 vs depth 3 / FAST
  for v in 32,41,31 32,43,31 32,44,31 32,34,31 32,141,31 32,134,31; do run $v --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --precision 2xf16; done
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
