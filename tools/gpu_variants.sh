#!/bin/bash
# Lean A/B of the aggregation kernel's variants.  FIELDCONV_B200_AGG_VARIANT=<b1 fp32>,<b1 packed>,<b2 fp32>,<b2 packed>;
# code = 100*FAST + 10*(CTAs/SM) + MODE (aggregate_kernel.cuh).  Parity of every variant first (tools/variant_probe.py).
#   gpurun --timeout 600 -- 'bash tools/gpu_variants.sh r02b'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run() {   # variant-string  layer_bench args...
  local v=$1; shift
  FIELDCONV_B200_AGG_VARIANT=$v timeout 100 python tools/layer_bench.py "$@" --tag "var$v"
}
{
  for v in 134 135 136 124 125 126 144 146; do
    for prec in 2xf16 2xf16p; do
      PROBE_PARITY_ONLY=1 PROBE_PRECISION=$prec FIELDCONV_B200_AGG_VARIANT=$v,$v,$v,$v timeout 100 python tools/variant_probe.py
    done
  done
} > $OUT/${TAG}_variant_parity.jsonl 2> $OUT/${TAG}_variant_parity.err
cat $OUT/${TAG}_variant_parity.jsonl | cut -c 1-400
{
  # band limit 2, fp32 output (the cfg-2 bench path)
  for v in 134 135 136 124 125 126 31; do run 0,0,$v,0 --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16; done
  # band limit 2, packed output
  for v in 22 134 135 136 124 125 126; do run 0,0,0,$v --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p; done
  # band limit 1, fp32 output (1 M vertices, C=32)
  for v in 134 135 136 144 146 41; do run $v,0,0,0 --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --precision 2xf16; done
  # band limit 1, packed output
  for v in 134 135 136 144 146; do run 0,$v,0,0 --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --precision 2xf16p; done
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
if [ -n "$NCU" ]; then
  FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate' -o $OUT/${TAG}_full_1m_c32 -f \
      python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --precision 2xf16 > $OUT/${TAG}_ncu_1m.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate' -o $OUT/${TAG}_full_cfg2 -f \
      python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16 > $OUT/${TAG}_ncu_cfg2.log 2>&1
  ls -la $OUT/${TAG}_full*
fi
