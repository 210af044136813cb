#!/bin/bash
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log | cut -c 1-300
{
  timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph --tag wide
  timeout 120 python tools/layer_bench.py --side 700 --channels 128 --band 1 --rings 6 --steps 5 --tag wide
  timeout 120 python tools/layer_bench.py --side 500 --channels 128 --band 2 --rings 6 --steps 5 --tag wide
  timeout 120 python tools/layer_bench.py --side 400 --channels 256 --band 1 --rings 6 --steps 5 --tag wide
} > $OUT/${TAG}_wide.jsonl 2> $OUT/${TAG}_wide.err
python - <<PY
import json
for l in open("$OUT/${TAG}_wide.jsonl"):
    d = json.loads(l)
    print(d["tag"], d["vertices"], d["channels"], d["band_limit"], d["n_rings"], "ms", d["ms_fwd_bwd"], d["kernels_ms"])
PY
tail -3 $OUT/${TAG}_wide.err
for lin in 2xf16 fp32; do
  FIELDCONV_B200_LIN=$lin timeout 300 python bench.py --skip-cfg4 --skip-cpu-baseline > $OUT/${TAG}_bench_lin_$lin.json 2> $OUT/${TAG}_bench_lin_$lin.err
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_lin_$lin.json").read().strip().splitlines()[-1])
print("lin=$lin", d["ms_per_step"], {k: v["ms"] for k, v in d["kernel_shares"].items() if k.startswith("lin_") or k in ("gemm_nn", "gemm_tn", "reduce_splits")})
PY
done
