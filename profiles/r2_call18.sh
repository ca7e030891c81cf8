set -x
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_b18_$tag.json 2> gpurun_out/r2_b18_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b18_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run default X=1
run w8 B200SCN_GPL_SMALL_WARPS=8
run w4 B200SCN_GPL_SMALL_WARPS=4
run t128_30k B200SCN_GPL_T128=30000
run t64_12k B200SCN_GPL_T64=12000
run mt1 B200SCN_GPL_SMALL_MT=1
run mt2 B200SCN_GPL_SMALL_MT=2
run default2 X=1
