set -x
timeout 600 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -k "large_input" 2>&1 | tail -12 > gpurun_out/r2_t14.log; tail -6 gpurun_out/r2_t14.log
run() { tag=$1; shift; env "$@" timeout 300 python profiles/subm_microbench.py > gpurun_out/r2_micro14_$tag.json 2> gpurun_out/r2_micro14_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r2_micro14_$tag.json')); print('$tag', round(d['launch_ms'],4), round(d['frac'],4), d.get('pair_list_layout'))"; tail -2 gpurun_out/r2_micro14_$tag.err; }
run d1_a1_r4 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=1 B200SCN_GPL_RING=4
run d1_a1_r6 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=1 B200SCN_GPL_RING=6
run d1_a0_r4 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=0 B200SCN_GPL_RING=4
run d0_a1_r4 B200SCN_GPL_DENSE=0 B200SCN_GPL_ACC128=1 B200SCN_GPL_RING=4
run d1_a1_r4_w10 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=1 B200SCN_GPL_RING=4 B200SCN_GPL_WARPS=10
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_gpl --launch-skip 3 -c 1 -o gpurun_out/r2_prof14_gpl -f python profiles/subm_microbench.py > gpurun_out/r2_ncu14.log 2>&1; tail -2 gpurun_out/r2_ncu14.log
