set -x
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((RANDOM % 10)) bench.py --gpus 2 --steps 40 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_n2d_$tag.json 2> gpurun_out/r2_n2d_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_n2d_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run default X=1
run min16 NCCL_MIN_CTAS=16
run min32 NCCL_MIN_CTAS=32
run nooverlap B200SCN_NO_OVERLAP_ALLREDUCE=1
run nooverlap_min32 B200SCN_NO_OVERLAP_ALLREDUCE=1 NCCL_MIN_CTAS=32
