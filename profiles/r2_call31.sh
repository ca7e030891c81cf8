set -x
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_b31_$tag.json 2> gpurun_out/r2_b31_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b31_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run default X=1
run bwd32 B200SCN_BN_FUSED_MAXG_BWD=32
run bwd96 B200SCN_BN_FUSED_MAXG_BWD=96
run bwd128 B200SCN_BN_FUSED_MAXG_BWD=128
