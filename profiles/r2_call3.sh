set -x
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bench_config.py tests/test_gpu_model.py tests/test_gpu_layers.py -q -m gpu --maxfail=12 -p no:cacheprovider -rA 2>&1 | grep -E "passed|failed|FAILED|Error|error|rows [0-9]+ loss|trajectory" | tail -40 > gpurun_out/r2_t3.log
tail -8 gpurun_out/r2_t3.log
timeout 300 python profiles/conv_levels_bench.py > gpurun_out/r2_levels_c2.log 2>&1; cat gpurun_out/r2_levels_c2.log
timeout 300 python profiles/conv_levels_bench.py --config c5 > gpurun_out/r2_levels_c5.log 2>&1; cat gpurun_out/r2_levels_c5.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; cut -c1-300 gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
B200SCN_TC_TWO=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench3_two0.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench3_two0.json
B200SCN_NO_FUSE_RES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench3_nores.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench3_nores.json
timeout 300 python profiles/step_timeline.py > gpurun_out/r2_timeline3.txt 2>&1; head -30 gpurun_out/r2_timeline3.txt
timeout 600 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench3_c5.json 2> gpurun_out/r2_bench3_c5.err; cut -c1-300 gpurun_out/r2_bench3_c5.json; tail -3 gpurun_out/r2_bench3_c5.err
