set -x
run() { tag=$1; cfg=$2; st=$3; shift; shift; shift; env "$@" timeout 300 python bench.py --config $cfg --steps $st --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_b28_$tag.json 2> gpurun_out/r2_b28_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b28_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run c3_t60k c3 20 B200SCN_GPL_T128=60000
run c3_t120k c3 20 B200SCN_GPL_T128=120000
run c3_t30k_t64_8k c3 20 B200SCN_GPL_T128=30000 B200SCN_GPL_T64=8000
run c5_t30k c5 6 B200SCN_GPL_T128=30000
run c5_t120k c5 6 B200SCN_GPL_T128=120000
