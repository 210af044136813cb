#!/bin/bash
# One gpurun call that regenerates the judged evidence for the current tree, most important first (a clamped call loses
# only the tail): parity tests, the driver's bench line (+ the reference arm), single-layer points (cfg 1 / cfg 3 as CUDA
# graphs, the cfg-2 layer, 1 M-vertex points, the cache-hostile permuted numbering), the ncu launch list of the bench
# command, `ncu --set full` of the layer kernels at the cfg-2 layer shape, at the real cfg-3 shape and at 1 M x C=32.
#   gpurun --timeout 1800 -- 'bash tools/gpu_evidence.sh r02p'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
timeout 600 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log | cut -c 1-200
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err | cut -c 1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
tail -c 300 $OUT/${TAG}_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
{
  timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph    # cfg 1 as a CUDA graph
  timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph   # cfg 3 as a CUDA graph
  timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6           # one cfg-2 layer
  timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --permute # ... cache-hostile numbering
  timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5          # cfg 5, HBM target point
  timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --permute
  timeout 120 python tools/layer_bench.py --side 700 --channels 128 --band 1 --rings 6 --steps 5          # tensor target point
  FIELDCONV_B200_FUSED=1 timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5 --tag fused_forward
} > $OUT/${TAG}_layers.jsonl 2> $OUT/${TAG}_layers.err
cut -c 1-420 $OUT/${TAG}_layers.jsonl; tail -3 $OUT/${TAG}_layers.err
kill $SMI
if [ -z "$SKIP_NCU" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
      --log-file $OUT/${TAG}_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 --skip-cfg4 --skip-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_h|k_pack_xhat' -o $OUT/${TAG}_full_cfg2 -f \
      python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 > $OUT/${TAG}_ncu_full_cfg2.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_h' -o $OUT/${TAG}_full_cfg3 -f \
      python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 > $OUT/${TAG}_ncu_full_cfg3.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_h' -o $OUT/${TAG}_full_1m_c32 -f \
      python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 > $OUT/${TAG}_ncu_full_1m_c32.log 2>&1
  ls -la $OUT/${TAG}_full* $OUT/${TAG}_ncu_launch_list.csv
fi
