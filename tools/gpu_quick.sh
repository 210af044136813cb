#!/bin/bash
# Short gpurun call while iterating on a kernel: the kernel tests and the parity tests in separate processes (a CUDA
# fault in one cannot poison the other), the default bench and two single-layer benches.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh r01d'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest_kernels.log 2>&1
echo "exit $?" >> $OUT/${TAG}_pytest_kernels.log
tail -15 $OUT/${TAG}_pytest_kernels.log
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partition.py -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest_parity.log 2>&1
echo "exit $?" >> $OUT/${TAG}_pytest_parity.log
tail -15 $OUT/${TAG}_pytest_parity.log
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1500 $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err
{
  timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6     # one cfg-2 layer
  timeout 120 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 --precision 3xtf32
  timeout 120 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6     # cfg 3
  timeout 120 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6    # cfg 5, HBM target point
  timeout 120 python tools/layer_bench.py --side 1000 --channels 128 --band 1 --rings 6   # cfg 5, tensor target point
} > $OUT/${TAG}_layers.jsonl 2> $OUT/${TAG}_layers.err
cat $OUT/${TAG}_layers.jsonl | cut -c 1-1200
tail -3 $OUT/${TAG}_layers.err
