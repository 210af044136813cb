#!/bin/bash
# One short gpurun call that regenerates the judged evidence for the current tree, most important first (a clamped
# call loses only the tail): parity tests, the driver's bench (+ the reference arm), `ncu --set full` of one layer at
# C=128 and at the cfg-2 layer shape, the ncu launch list of the bench command, two large single-layer points.
#   gpurun --timeout 900 -- 'bash tools/gpu_evidence.sh r01e'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
if [ -z "$SKIP_NCU" ]; then
FIELDCONV_B200_NCU=1 timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_c128 -f \
    python tools/layer_bench.py --side 284 --channels 128 --band 1 --rings 6 > $OUT/${TAG}_ncu_full_c128.log 2>&1
FIELDCONV_B200_NCU=1 timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_cfg2 -f \
    python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 > $OUT/${TAG}_ncu_full_cfg2.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file $OUT/${TAG}_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu_bench.log 2>&1
fi
{
  timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph    # cfg 1 as a CUDA graph
  timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph   # cfg 3 as a CUDA graph
  timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6           # one cfg-2 layer
  timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6          # cfg 5, HBM target point
  timeout 120 python tools/layer_bench.py --side 1000 --channels 128 --band 1 --rings 6         # cfg 5, tensor target point
} > $OUT/${TAG}_layers.jsonl 2> $OUT/${TAG}_layers.err
cut -c 1-700 $OUT/${TAG}_layers.jsonl; tail -3 $OUT/${TAG}_layers.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench_ref.json
ls -la $OUT | tail -12
