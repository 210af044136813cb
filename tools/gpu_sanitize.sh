#!/bin/bash
# compute-sanitizer passes over the GPU tests of the kernels added or rewritten in round 2 (memcheck on all of them,
# racecheck + synccheck on the shared-memory-heavy ones).  Writes gpurun_out/<tag>_sanitize_*.log.
#   gpurun --timeout 2400 -- 'bash tools/gpu_sanitize.sh r02t'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
CS="compute-sanitizer --error-exitcode 9 --launch-timeout 0 --target-processes all"
run() {   # name tool timeout tests...
  local name=$1 tool=$2 tmo=$3; shift 3
  timeout $tmo $CS --tool $tool python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > $OUT/${TAG}_sanitize_${name}.log 2>&1
  echo "$name ($tool): exit $?  $(grep -c 'ERROR SUMMARY' $OUT/${TAG}_sanitize_${name}.log) summaries: $(grep 'ERROR SUMMARY' $OUT/${TAG}_sanitize_${name}.log | sort | uniq -c | tr '\n' ';')  $(tail -1 $OUT/${TAG}_sanitize_${name}.log)"
}
run mem_new memcheck 900 tests/test_gpu_fused.py tests/test_gpu_lift.py tests/test_gpu_echo.py tests/test_gpu_support_graph.py
run mem_kernels memcheck 900 tests/test_gpu_kernels.py tests/test_gpu_packed.py
run race_new racecheck 600 tests/test_gpu_echo.py tests/test_gpu_lift.py
run race_kernels racecheck 900 tests/test_gpu_kernels.py tests/test_gpu_parity.py -k "aggregate_matches or degree_skewed or compact_plan_path or modrelu"
run mem_parity memcheck 900 tests/test_gpu_parity.py -k "degree_skewed or block_epilogue or weight_gradient_from_g or compact_plan_path or edge_cases"
run sync_fused synccheck 600 tests/test_gpu_fused.py
