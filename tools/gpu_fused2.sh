#!/bin/bash
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log | cut -c 1-300
{
  export FIELDCONV_B200_FUSED=1
  timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph --tag "fused=1"
  timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --tag "fused=1"
  timeout 120 python tools/layer_bench.py --side 1000 --channels 64 --band 1 --rings 6 --deg 64 --steps 5 --tag "fused=1"
} > $OUT/${TAG}_fused_ab.jsonl 2> $OUT/${TAG}_fused_ab.err
python - <<PY
import json
for l in open("$OUT/${TAG}_fused_ab.jsonl"):
    d = json.loads(l)
    print(d["tag"], d["vertices"], d["channels"], d["band_limit"], d["n_rings"], d["precision"], "ms", d["ms_fwd_bwd"], d["kernels_ms"])
PY
tail -3 $OUT/${TAG}_fused_ab.err
