set -x
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bench_config.py -q -m gpu --maxfail=12 -p no:cacheprovider -rA 2>&1 | grep -E "passed|failed|FAILED|Error|error|rows [0-9]+ loss|trajectory" | tail -40 > gpurun_out/r2_t4.log
tail -4 gpurun_out/r2_t4.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; cut -c1-300 gpurun_out/r2_bench4.json; tail -5 gpurun_out/r2_bench4.err
B200SCN_TC_TWO=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench4_two0.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench4_two0.json
timeout 300 python profiles/step_timeline.py > gpurun_out/r2_timeline4.txt 2>&1; head -30 gpurun_out/r2_timeline4.txt
timeout 600 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench4_c5.json 2> gpurun_out/r2_bench4_c5.err; cut -c1-300 gpurun_out/r2_bench4_c5.json; tail -3 gpurun_out/r2_bench4_c5.err
timeout 300 python profiles/conv_levels_bench.py --config c5 > gpurun_out/r2_levels_c5.log 2>&1; cat gpurun_out/r2_levels_c5.log
