#!/bin/bash
# A/B of the fused forward (FIELDCONV_B200_FUSED=0/1) on single layers + one ncu --set full capture of the fused kernel.
#   gpurun --timeout 900 -- 'bash tools/gpu_fused.sh r02g'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
{
  for v in 0 1; do
    export FIELDCONV_B200_FUSED=$v
    timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph --tag "fused=$v"
    timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --tag "fused=$v"
    timeout 120 python tools/layer_bench.py --side 1000 --channels 64 --band 1 --rings 6 --deg 64 --steps 5 --tag "fused=$v"
    timeout 120 python tools/layer_bench.py --side 700 --channels 128 --band 1 --rings 6 --steps 5 --tag "fused=$v"
    timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 2 --steps 5 --tag "fused=$v"
  done
} > $OUT/${TAG}_fused_ab.jsonl 2> $OUT/${TAG}_fused_ab.err
python - <<PY
import json
for l in open("$OUT/${TAG}_fused_ab.jsonl"):
    d = json.loads(l)
    print(d["tag"], d["vertices"], d["channels"], d["band_limit"], d["n_rings"], d["precision"], "ms", d["ms_fwd_bwd"], d["kernels_ms"])
PY
tail -3 $OUT/${TAG}_fused_ab.err
FIELDCONV_B200_FUSED=1 FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_fused_fwd' -o $OUT/${TAG}_full_fused_1m_c32 -f \
    python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 > $OUT/${TAG}_ncu_fused.log 2>&1
FIELDCONV_B200_FUSED=1 FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_fused_fwd' -o $OUT/${TAG}_full_fused_cfg1 -f \
    python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 > $OUT/${TAG}_ncu_fused_cfg1.log 2>&1
ls -la $OUT/${TAG}_full*
