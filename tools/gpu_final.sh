#!/bin/bash
# Round-end evidence for the tree as committed: parity suite, the driver's bench line, ncu --set full of one cfg-2 layer,
# the ncu launch list of the bench command, single-layer points.  Most important first.
#   gpurun --timeout 480 -- 'bash tools/gpu_final.sh r01j'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests -m gpu -q -rf --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log | cut -c 1-200
timeout 200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cut -c 1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
{
  timeout 60 python tools/layer_bench.py --side 71 --channels 32 --band 1 --rings 6 --graph    # cfg 1 as a CUDA graph
  timeout 60 python tools/layer_bench.py --side 83 --channels 128 --band 2 --rings 6 --graph   # cfg 3 as a CUDA graph
  timeout 60 python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6           # one cfg-2 layer
  timeout 90 python tools/layer_bench.py --side 1000 --channels 32 --band 1 --rings 6 --steps 5   # cfg 5, HBM target point
  timeout 90 python tools/layer_bench.py --side 1000 --channels 128 --band 1 --rings 6 --steps 3  # cfg 5, tensor target point
} > $OUT/${TAG}_layers.jsonl 2> $OUT/${TAG}_layers.err
cut -c 1-420 $OUT/${TAG}_layers.jsonl; tail -3 $OUT/${TAG}_layers.err
FIELDCONV_B200_NCU=1 timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_aggregate|k_gemm_tc|k_gemm_h' -o $OUT/${TAG}_full_cfg2 -f \
    python tools/layer_bench.py --side 284 --channels 48 --band 2 --rings 6 > $OUT/${TAG}_ncu_full_cfg2.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file $OUT/${TAG}_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
ls -la $OUT | grep ${TAG}
