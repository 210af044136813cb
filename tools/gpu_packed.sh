#!/bin/bash
# A/B call for the packed-operand path (precision="2xf16p") and the occupancy-3 aggregation variant:
# packed parity tests first (own process, own timeout: a hung kernel must not eat the call), then benches of both paths,
# ncu captures of the packed kernels last.
#   gpurun --timeout 900 -- 'bash tools/gpu_packed.sh r01f'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
FIELDCONV_B200_TEST_PACKED=1 timeout 300 python -m pytest tests/test_gpu_packed.py -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest_packed.log 2>&1
PK_RC=$?
echo "pytest exit $PK_RC" >> $OUT/${TAG}_pytest_packed.log
tail -25 $OUT/${TAG}_pytest_packed.log | cut -c 1-300
timeout 300 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log | cut -c 1-300
timeout 240 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1800 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
{
  FIELDCONV_B200_AGG_OCC=3 timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --tag occ3
  FIELDCONV_B200_AGG_OCC=3 timeout 120 python tools/layer_bench.py --side 284 --channels 128 --band 1 --rings 6 --tag occ3
  timeout 120 python tools/layer_bench.py --side 284 --channels 128 --band 1 --rings 6
} > $OUT/${TAG}_layers_occ3.jsonl 2> $OUT/${TAG}_layers_occ3.err
cut -c 1-600 $OUT/${TAG}_layers_occ3.jsonl; tail -3 $OUT/${TAG}_layers_occ3.err
if [ $PK_RC -ne 124 ]; then
  timeout 240 python bench.py --precision 2xf16p > $OUT/${TAG}_bench_packed.json 2> $OUT/${TAG}_bench_packed.err
  tail -c 1800 $OUT/${TAG}_bench_packed.json; tail -3 $OUT/${TAG}_bench_packed.err
  {
    timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p
    timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --precision 2xf16p
    timeout 120 python tools/layer_bench.py --side 1000 --channels 128 --band 1 --rings 6 --precision 2xf16p
    timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --precision 2xf16p --graph
    timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --precision 2xf16p --graph
    FIELDCONV_B200_AGG_OCC=3 timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --precision 2xf16p --tag occ3
  } > $OUT/${TAG}_layers_packed.jsonl 2> $OUT/${TAG}_layers_packed.err
  cut -c 1-700 $OUT/${TAG}_layers_packed.jsonl; tail -3 $OUT/${TAG}_layers_packed.err
  FIELDCONV_B200_NCU=1 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_cfg2_packed -f \
      python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --precision 2xf16p > $OUT/${TAG}_ncu_full_cfg2_packed.log 2>&1
  FIELDCONV_B200_NCU=1 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_c128_packed -f \
      python tools/layer_bench.py --side 284 --channels 128 --band 1 --rings 6 --precision 2xf16p > $OUT/${TAG}_ncu_full_c128_packed.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
      --log-file $OUT/${TAG}_ncu_launch_list_packed.csv python bench.py --precision 2xf16p --steps 2 --warmup 1 > $OUT/${TAG}_ncu_bench_packed.log 2>&1
fi
ls -la $OUT | tail -14
