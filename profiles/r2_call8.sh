set -x
timeout 300 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t8a.log 2>&1; OK=$?
tail -6 gpurun_out/r2_t8a.log
if [ $OK -ne 0 ]; then echo "MMA TESTS FAILED"; exit 1; fi
for knobs in "B200SCN_TC_WIDE=1" "B200SCN_TC_WIDE=0" "B200SCN_TC_MIN_C=64 B200SCN_TC_WGRAD_MIN_C=64" "B200SCN_TC_MIN_C=64"; do
  tag=$(echo $knobs | tr ' =' '__')
  env $knobs timeout 300 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench8_c5_$tag.json 2>/dev/null; echo $tag; cut -c1-180 gpurun_out/r2_bench8_c5_$tag.json
done
timeout 300 python profiles/step_timeline.py --config c5 --steps 2 > gpurun_out/r2_timeline8_c5.txt 2>&1; head -16 gpurun_out/r2_timeline8_c5.txt | tail -12
timeout 600 python -m pytest tests/test_gpu_bench_config.py -q -m gpu -p no:cacheprovider -rA -k "c5" 2>&1 | grep -E "passed|failed|FAILED|rows [0-9]+ loss" | tail -8
