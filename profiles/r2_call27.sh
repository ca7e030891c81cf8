set -x
run() { tag=$1; cfg=$2; shift; shift; env "$@" timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_b27_$tag.json 2> gpurun_out/r2_b27_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b27_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run c2_default c2 X=1
run c2_t30k c2 B200SCN_GPL_T128=30000
run c2_t30k_mt2 c2 B200SCN_GPL_T128=30000 B200SCN_GPL_SMALL_MT=2
run c3_default c3 X=1
run c3_t30k c3 B200SCN_GPL_T128=30000
run c1_default c1 X=1
run c1_t30k c1 B200SCN_GPL_T128=30000
