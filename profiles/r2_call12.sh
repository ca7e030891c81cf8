set -x
timeout 600 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -k "large_input" -x 2>&1 | tail -4 > gpurun_out/r2_t12.log; tail -4 gpurun_out/r2_t12.log
run() { tag=$1; shift; env "$@" timeout 300 python profiles/subm_microbench.py > gpurun_out/r2_micro12_$tag.json 2> gpurun_out/r2_micro12_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r2_micro12_$tag.json')); print('$tag', round(d['launch_ms'],4), round(d['frac'],4), d.get('pair_list_layout'))"; }
run d1_a1_pf0 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=1 B200SCN_GPL_PF=0
run d1_a0_pf0 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=0 B200SCN_GPL_PF=0
run d1_a1_pf3 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=1 B200SCN_GPL_PF=3
run d1_a0_pf3 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=0 B200SCN_GPL_PF=3
run d1_a0_pf2 B200SCN_GPL_DENSE=1 B200SCN_GPL_ACC128=0 B200SCN_GPL_PF=2
run d0_a0_pf0 B200SCN_GPL_DENSE=0 B200SCN_GPL_ACC128=0 B200SCN_GPL_PF=0
run d3_a1_pf0 B200SCN_GPL_DENSE=3 B200SCN_GPL_ACC128=1 B200SCN_GPL_PF=0
timeout 400 python bench.py --config c5 --steps 4 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench12_c5.json 2> gpurun_out/r2_bench12_c5.err; cut -c1-200 gpurun_out/r2_bench12_c5.json; tail -3 gpurun_out/r2_bench12_c5.err
