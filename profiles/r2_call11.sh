set -x
timeout 600 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -k "large_input or gpl_and_pairs or gpl_lists" -x 2>&1 | tail -8 > gpurun_out/r2_t11.log; tail -8 gpurun_out/r2_t11.log
for d in 0 1 3; do
  B200SCN_GPL_DENSE=$d timeout 300 python profiles/subm_microbench.py > gpurun_out/r2_micro11_d$d.json 2> gpurun_out/r2_micro11_d$d.err; python -c "
import json; d=json.load(open('gpurun_out/r2_micro11_d$d.json')); print('dense', $d, d['launch_ms'], d['frac'], d.get('pair_list_layout'))"; tail -2 gpurun_out/r2_micro11_d$d.err
done
B200SCN_GPL_DENSE=3 timeout 300 python profiles/subm_microbench.py --math tf32 > gpurun_out/r2_micro11_d3_tf32.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_micro11_d3_tf32.json')); print('dense 3 tf32', d['launch_ms'], d['frac'])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_gpl --launch-skip 3 -c 1 -o gpurun_out/r2_prof11_gpl -f python profiles/subm_microbench.py > gpurun_out/r2_ncu11.log 2>&1; tail -2 gpurun_out/r2_ncu11.log
