#!/bin/bash
# One gpurun call: GPU parity tests, the driver's bench line (N = 1, with the cfg-4 sub-record) and the reference arm.
#   gpurun --timeout 1200 -- 'bash tools/gpu_bench.sh r02d'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
T0=$(date +%s); timeout 900 python bench.py ${BENCH_ARGS} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $? wall $(( $(date +%s) - T0 )) s"; tail -5 $OUT/${TAG}_bench.err | cut -c 1-600
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"])
    r = d["roofline"]; print({k: r[k] for k in r if k not in ("per_kernel", "note")})
    print(json.dumps(r.get("per_kernel"))[:1500])
    print("cfg4", json.dumps(d.get("cfg4"))[:1800])
    print("cpu", d.get("cpu_baseline"))
    print(json.dumps(d["kernel_shares"])[:2500])
except Exception as e:
    print("parse failed", e)
PY
if [ -z "$SKIP_REF" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
  cat $OUT/${TAG}_bench_ref.json | cut -c 1-1500; tail -3 $OUT/${TAG}_bench_ref.err
fi
