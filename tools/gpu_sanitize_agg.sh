OUT=gpurun_out; mkdir -p $OUT
CS="compute-sanitizer --error-exitcode 9 --launch-timeout 0 --target-processes all"
timeout 120 $CS --tool memcheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_packed.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "aggregate or degree_skewed or compact_plan_path or packed_scale or decode or edge_cases" > $OUT/r04k_sanitize_mem.log 2>&1
echo "memcheck exit $? $(grep 'ERROR SUMMARY' $OUT/r04k_sanitize_mem.log | sort | uniq -c | tr '\n' ';') $(tail -1 $OUT/r04k_sanitize_mem.log)"
timeout 100 $CS --tool racecheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "aggregate_matches or degree_skewed" > $OUT/r04k_sanitize_race.log 2>&1
echo "racecheck exit $? $(grep 'RACECHECK SUMMARY' $OUT/r04k_sanitize_race.log | sort | uniq -c | tr '\n' ';') $(tail -1 $OUT/r04k_sanitize_race.log)"
