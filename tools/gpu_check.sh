#!/bin/bash
# One gpurun call: GPU parity tests, then single-layer timings (optionally A/B over an environment switch).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh r02c FIELDCONV_B200_AGG_CM2 0 1'
TAG=${1:-rXX}; shift
VAR=${1:-NONE}; shift
VALS=${@:-0}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
{
  for v in $VALS; do
    export $VAR=$v
    timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --tag "$VAR=$v"
    timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p --tag "$VAR=$v"
    timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --tag "$VAR=$v"
    timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph --tag "$VAR=$v"
    timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph --tag "$VAR=$v"
  done
} > $OUT/${TAG}_layers.jsonl 2> $OUT/${TAG}_layers.err
python - <<PY
import json
for l in open("$OUT/${TAG}_layers.jsonl"):
    d = json.loads(l)
    print(d["tag"], d["vertices"], d["channels"], d["band_limit"], d["precision"], "ms", d["ms_fwd_bwd"], d["kernels_ms"])
PY
tail -3 $OUT/${TAG}_layers.err
