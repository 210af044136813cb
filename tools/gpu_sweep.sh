#!/bin/bash
# pytest + bench (N = 1) + the cfg-5 sweep (BASELINE configs[4]) on one GPU.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sweep.sh r02j'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SKIP_REF=1 bash tools/gpu_bench.sh $TAG
if [ -z "$SKIP_SWEEP" ]; then
  timeout 900 python tools/layer_bench.py --sweep --steps 6 > $OUT/${TAG}_sweep.jsonl 2> $OUT/${TAG}_sweep.err
  python - <<PY
import json
for l in open("$OUT/${TAG}_sweep.jsonl"):
    d = json.loads(l)
    if "skipped" in d: print("skipped", d["skipped"], d["why"]); continue
    k = d["kernels_ms"]
    print("C=%d B=%d R=%d  %.2f ms  %.2f Gedges/s  hbm_frac %.3f  flags %s" % (d["channels"], d["band_limit"], d["n_rings"], d["ms_fwd_bwd"], d["edges_per_s"] / 1e9, d["hbm_frac"], d["flags"]),
          {n: round(v, 2) for n, v in sorted(k.items(), key=lambda kv: -kv[1])[:5]})
PY
  tail -3 $OUT/${TAG}_sweep.err
fi
