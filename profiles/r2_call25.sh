set -x
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_bench_config.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2_t25.log; cat gpurun_out/r2_t25.log
timeout 400 python bench.py --config c5 --steps 6 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench25_c5.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench25_c5.json
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench25_c3.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench25_c3.json
timeout 300 python profiles/step_timeline.py --config c5 --steps 2 > gpurun_out/r2_timeline25_c5.txt 2>&1; head -22 gpurun_out/r2_timeline25_c5.txt | tail -19
