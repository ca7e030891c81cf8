set -x
timeout 600 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -k "large_input" 2>&1 | tail -12 > gpurun_out/r2_t15.log; tail -6 gpurun_out/r2_t15.log
run() { tag=$1; shift; env "$@" timeout 300 python profiles/subm_microbench.py > gpurun_out/r2_micro15_$tag.json 2> gpurun_out/r2_micro15_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r2_micro15_$tag.json')); print('$tag', round(d['launch_ms'],4), round(d['frac'],4), d.get('pair_list_layout'))"; tail -2 gpurun_out/r2_micro15_$tag.err; }
run d1_b1 B200SCN_GPL_DENSE=1 B200SCN_GPL_BATCH=1
run d1_b2 B200SCN_GPL_DENSE=1 B200SCN_GPL_BATCH=2
run d1_b3 B200SCN_GPL_DENSE=1 B200SCN_GPL_BATCH=3
run d0_b2 B200SCN_GPL_DENSE=0 B200SCN_GPL_BATCH=2
run d1_b2_w16 B200SCN_GPL_DENSE=1 B200SCN_GPL_BATCH=2 B200SCN_GPL_WARPS=16
run d1_b2_a0 B200SCN_GPL_DENSE=1 B200SCN_GPL_BATCH=2 B200SCN_GPL_ACC128=0
