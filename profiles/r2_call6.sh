set -x
timeout 150 python -m pytest tests/test_gpu_mma.py -q -m gpu -x -p no:cacheprovider -k "wgrad_tc" > gpurun_out/r2_t6a.log 2>&1; WG=$?
tail -15 gpurun_out/r2_t6a.log
if [ $WG -ne 0 ]; then export B200SCN_NO_TC_WGRAD=1; echo "WGRAD TC DISABLED"; fi
for knobs in "B200SCN_TC_DRAIN=0 B200SCN_TC_NACC=1" "B200SCN_TC_DRAIN=1 B200SCN_TC_NACC=1" "B200SCN_TC_DRAIN=0 B200SCN_TC_NACC=4" "B200SCN_TC_DRAIN=1 B200SCN_TC_NACC=4"; do
  tag=$(echo $knobs | tr ' =' '__')
  env $knobs B200SCN_NO_TC_WGRAD=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench6_$tag.json 2>/dev/null; echo $tag; cut -c1-160 gpurun_out/r2_bench6_$tag.json
done
B200SCN_NO_TC_WGRAD=1 timeout 200 python profiles/tc_micro.py --level 3 --reps 2 > gpurun_out/r2_tcmicro_plain.log 2>&1 && B200SCN_NO_TC_WGRAD=1 timeout 420 ncu --set full --import-source on --clock-control none -k regex:k_sparse_conv_tc -c 2 -o gpurun_out/prof_tc6 python profiles/tc_micro.py --level 3 --reps 2 > gpurun_out/r2_ncu6.log 2>&1
tail -3 gpurun_out/r2_ncu6.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; cut -c1-200 gpurun_out/r2_bench6.json; tail -5 gpurun_out/r2_bench6.err
timeout 200 python profiles/step_timeline.py > gpurun_out/r2_timeline6.txt 2>&1; head -24 gpurun_out/r2_timeline6.txt | tail -20
