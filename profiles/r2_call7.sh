set -x
timeout 150 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -k "wgrad_tc" > gpurun_out/r2_t7a.log 2>&1; WG=$?
tail -12 gpurun_out/r2_t7a.log
if [ $WG -ne 0 ]; then export B200SCN_NO_TC_WGRAD=1; echo "WGRAD TC DISABLED"; fi
timeout 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bench_config.py tests/test_gpu_model.py -q -m gpu --maxfail=8 -p no:cacheprovider -rA --deselect tests/test_gpu_mma.py::test_wgrad_tc_matches_ffma_kernel 2>&1 | grep -E "passed|failed|FAILED|Error|error|rows [0-9]+ loss|trajectory" | tail -40 > gpurun_out/r2_t7.log
tail -8 gpurun_out/r2_t7.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; cut -c1-200 gpurun_out/r2_bench7.json; tail -5 gpurun_out/r2_bench7.err
timeout 200 python profiles/step_timeline.py > gpurun_out/r2_timeline7.txt 2>&1; head -24 gpurun_out/r2_timeline7.txt | tail -20
timeout 400 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench7_c5.json 2> gpurun_out/r2_bench7_c5.err; cut -c1-200 gpurun_out/r2_bench7_c5.json; tail -3 gpurun_out/r2_bench7_c5.err
timeout 300 python profiles/step_timeline.py --config c5 --steps 2 > gpurun_out/r2_timeline7_c5.txt 2>&1; head -24 gpurun_out/r2_timeline7_c5.txt | tail -20
