set -x
timeout 600 python -m pytest tests/test_gpu_mma.py -q -m gpu --maxfail=6 -p no:cacheprovider -k "wgrad_tc" 2>&1 | tail -30 > gpurun_out/r2_t5a.log
tail -12 gpurun_out/r2_t5a.log
for n in 1 2 4; do
B200SCN_NO_TC_WGRAD=1 B200SCN_TC_NACC=$n timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench5_nacc$n.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench5_nacc$n.json
done
B200SCN_NO_TC_WGRAD=1 timeout 300 python profiles/step_timeline.py > gpurun_out/r2_timeline5.txt 2>&1; head -16 gpurun_out/r2_timeline5.txt | tail -12
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_bench_config.py -q -m gpu --maxfail=12 -p no:cacheprovider -rA 2>&1 | grep -E "passed|failed|FAILED|Error|error|rows [0-9]+ loss|trajectory" | tail -40 > gpurun_out/r2_t5.log
tail -6 gpurun_out/r2_t5.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; cut -c1-200 gpurun_out/r2_bench5.json; tail -5 gpurun_out/r2_bench5.err
timeout 300 python profiles/step_timeline.py > gpurun_out/r2_timeline5b.txt 2>&1; head -20 gpurun_out/r2_timeline5b.txt | tail -16
timeout 600 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench5_c5.json 2> gpurun_out/r2_bench5_c5.err; cut -c1-200 gpurun_out/r2_bench5_c5.json; tail -3 gpurun_out/r2_bench5_c5.err
