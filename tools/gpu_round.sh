#!/bin/bash
# One gpurun call: parity tests, the driver's bench, single-layer benches (cfg 1/3/5 points), the ncu launch
# list of the bench command and one `--set full` capture of a single fwd+bwd layer step.  Everything lands in
# gpurun_out/<tag>_*; summaries are copied to profiles/ by hand afterwards.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01c'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
if [ -z "$SKIP_TESTS" ]; then
  timeout 600 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json
{
  timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6      # cfg 1
  timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6     # cfg 3
  timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6     # one cfg-2 layer
  timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6    # cfg 5, HBM target point
  timeout 120 python tools/layer_bench.py --side 1000 --channels 128 --band 1 --rings 6   # cfg 5, tensor target point
  timeout 120 python tools/layer_bench.py --side 1000 --channels 64 --band 1 --rings 6 --deg 64   # a cfg-4 rank's shape
} > $OUT/${TAG}_layers.jsonl 2> $OUT/${TAG}_layers.err
cat $OUT/${TAG}_layers.jsonl | cut -c 1-900
kill $SMI
if [ -z "$SKIP_NCU" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
      --log-file $OUT/${TAG}_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu_bench.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_cfg2 -f \
      python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 > $OUT/${TAG}_ncu_full_cfg2.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_c128 -f \
      python tools/layer_bench.py --side 284 --channels 128 --band 1 --rings 6 > $OUT/${TAG}_ncu_full_c128.log 2>&1
  ls -la $OUT | tail -12
fi
