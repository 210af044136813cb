#!/bin/bash
# gpurun call for the ncu evidence only: `--set full` captures of one FieldConv layer fwd+bwd at the cfg-2 layer shape
# and at C=128, plus (optional, FAST=1 skips) the tests and the bench.
#   gpurun --timeout 1200 -- 'bash tools/gpu_ncu.sh r01f'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
if [ -z "$FAST" ]; then
  timeout 600 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -15 $OUT/${TAG}_pytest.log
  timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  tail -c 1200 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
  {
    timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6            # cfg 1 eager
    timeout 120 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph    # cfg 1 as a CUDA graph
    timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph   # cfg 3 as a CUDA graph
  } > $OUT/${TAG}_layers_graph.jsonl 2> $OUT/${TAG}_layers_graph.err
  cut -c 1-400 $OUT/${TAG}_layers_graph.jsonl; tail -3 $OUT/${TAG}_layers_graph.err
fi
FIELDCONV_B200_NCU=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_cfg2 -f \
    python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 > $OUT/${TAG}_ncu_full_cfg2.log 2>&1
FIELDCONV_B200_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_c128 -f \
    python tools/layer_bench.py --side 284 --channels 128 --band 1 --rings 6 > $OUT/${TAG}_ncu_full_c128.log 2>&1
ls -la $OUT | tail -8
