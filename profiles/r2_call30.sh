set -x
run() { tag=$1; cfg=$2; st=$3; shift; shift; shift; env "$@" timeout 300 python bench.py --config $cfg --steps $st --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_b30_$tag.json 2> gpurun_out/r2_b30_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b30_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run c2_default c2 40 X=1
run c2_t64_3000 c2 40 B200SCN_GPL_T64=3000
run c2_t64_1000 c2 40 B200SCN_GPL_T64=1000
run c3_t64_3000 c3 20 B200SCN_GPL_T64=3000
