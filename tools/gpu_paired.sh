#!/bin/bash
# Lean A/B of the paired-chunk producer variant of k_gemm_h_nn (FIELDCONV_B200_GEMM_PAIRED=1).
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
FIELDCONV_B200_GEMM_PAIRED=1 timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest_paired.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_paired.log
tail -4 $OUT/${TAG}_pytest_paired.log | cut -c 1-200
{
  timeout 100 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --tag base
  FIELDCONV_B200_GEMM_PAIRED=1 timeout 100 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --tag paired
  timeout 100 python tools/layer_bench.py --side 284 --channels 128 --band 2 --rings 6 --tag base
  FIELDCONV_B200_GEMM_PAIRED=1 timeout 100 python tools/layer_bench.py --side 284 --channels 128 --band 2 --rings 6 --tag paired
  timeout 100 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --tag base
  FIELDCONV_B200_GEMM_PAIRED=1 timeout 100 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --tag paired
} > $OUT/${TAG}_paired.jsonl 2> $OUT/${TAG}_paired.err
python - <<PY
import json
for l in open("$OUT/${TAG}_paired.jsonl"):
    d = json.loads(l)
    k = d["kernels_ms"]
    print(d["tag"], d["vertices"], d["channels"], d["band_limit"], d["precision"], "ms", d["ms_fwd_bwd"],
          {n: v for n, v in k.items() if n.startswith(("aggregate", "gemm"))})
PY
tail -3 $OUT/${TAG}_paired.err
