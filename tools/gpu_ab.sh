#!/bin/bash
# Lean A/B call: full parity suite with the current defaults, then the cfg-2 bench under the candidate settings.
#   gpurun --timeout 600 -- 'bash tools/gpu_ab.sh r01g'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log | cut -c 1-300
timeout 200 python bench.py --skip-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
FIELDCONV_B200_AGG_OCC3_MAX_B=2 timeout 200 python bench.py --skip-cpu-baseline > $OUT/${TAG}_bench_occ3b2.json 2> $OUT/${TAG}_bench_occ3b2.err
timeout 200 python bench.py --skip-cpu-baseline --precision 2xf16p > $OUT/${TAG}_bench_packed.json 2> $OUT/${TAG}_bench_packed.err
FIELDCONV_B200_AGG_OCC3_MAX_B=2 timeout 200 python bench.py --skip-cpu-baseline --precision 2xf16p > $OUT/${TAG}_bench_packed_occ3b2.json 2> $OUT/${TAG}_bench_packed_occ3b2.err
for f in default occ3b2 packed packed_occ3b2; do echo "== $f"; cut -c 1-260 $OUT/${TAG}_bench_$f.json; tail -2 $OUT/${TAG}_bench_$f.err; done
{
  FIELDCONV_B200_AGG_OCC3_MAX_B=2 timeout 100 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --tag occ3b2
  FIELDCONV_B200_AGG_OCC3_MAX_B=2 timeout 100 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p --tag occ3b2
  FIELDCONV_B200_AGG_OCC3_MAX_B=2 timeout 100 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph --tag occ3b2
} > $OUT/${TAG}_layers_ab.jsonl 2> $OUT/${TAG}_layers_ab.err
cut -c 1-600 $OUT/${TAG}_layers_ab.jsonl; tail -3 $OUT/${TAG}_layers_ab.err
