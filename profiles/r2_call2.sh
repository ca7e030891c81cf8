set -x
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bench_config.py tests/test_gpu_model.py -q -m gpu --maxfail=12 -p no:cacheprovider -rA 2>&1 | grep -E "passed|failed|FAILED|Error|error|rows [0-9]+ loss|trajectory" | tail -60 > gpurun_out/r2_t2.log
tail -25 gpurun_out/r2_t2.log
timeout 300 python profiles/conv_levels_bench.py > gpurun_out/r2_levels_c2.log 2>&1; cat gpurun_out/r2_levels_c2.log
timeout 300 python profiles/conv_levels_bench.py --config c5 > gpurun_out/r2_levels_c5.log 2>&1; cat gpurun_out/r2_levels_c5.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; cut -c1-400 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
B200SCN_NO_TC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench2_notc.json 2> gpurun_out/r2_bench2_notc.err; cut -c1-300 gpurun_out/r2_bench2_notc.json
timeout 300 python profiles/step_timeline.py > gpurun_out/r2_timeline2.txt 2>&1; head -40 gpurun_out/r2_timeline2.txt
timeout 600 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench2_c5.json 2> gpurun_out/r2_bench2_c5.err; cut -c1-300 gpurun_out/r2_bench2_c5.json; tail -3 gpurun_out/r2_bench2_c5.err
