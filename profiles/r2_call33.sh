set -x
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_b33_$tag.json 2> gpurun_out/r2_b33_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_b33_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run fwd64 X=1
run fwd48 B200SCN_BN_FUSED_MAXG=48 B200SCN_BN_FUSED_MAXG_BWD=96
run fwd96 B200SCN_BN_FUSED_MAXG=96 B200SCN_BN_FUSED_MAXG_BWD=96
BEST=$(python - <<'PY'
import json
best=None
for t,v in (("fwd64",64),("fwd48",48),("fwd96",96)):
    d=json.loads(open("gpurun_out/r2_b33_%s.json"%t).read().strip().splitlines()[-1])
    if best is None or d["ms_per_step"]<best[0]-0.01: best=(d["ms_per_step"],v)
print(best[1])
PY
)
echo BEST=$BEST
B200SCN_BN_FUSED_MAXG=$BEST B200SCN_BN_FUSED_MAXG_BWD=96 timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_bench_config.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
