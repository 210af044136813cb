#!/bin/bash
# Multi-GPU evidence (gpurun --gpus N): the 2-rank NCCL partition test, then the driver's bench command at N ranks
# (data-parallel cfg-2 line with the NCCL partition parity and the cfg-4 sub-record inside).
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_multi.sh r02i 2'
TAG=${1:-rXX}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ -z "$SKIP_TESTS" ]; then
  timeout 600 python -m pytest tests -m gpu -q -rs --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest_partition_n2.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_partition_n2.log
  tail -5 $OUT/${TAG}_pytest_partition_n2.log | cut -c 1-300
fi
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench exit $? wall $(( $(date +%s) - T0 )) s"; tail -4 $OUT/${TAG}_bench_n$N.err | cut -c 1-800
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("parity", json.dumps(d.get("partition_parity")))
    c = d.get("cfg4") or {}
    print("cfg4", {k: c.get(k) for k in ("value", "ms_per_step", "n_gpus", "setup_s")}, json.dumps(c.get("halo")), json.dumps(c.get("kernel_ms_rank0"))[:900])
except Exception as e:
    print("parse failed", e)
PY
