set -x
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$((RANDOM % 10)) bench.py --gpus 2 --steps 40 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_n2c_$tag.json 2> gpurun_out/r2_n2c_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_n2c_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), '%.4e' % d['value'])"; }
run default X=1
run ctas4 NCCL_MAX_CTAS=4
run ctas2 NCCL_MAX_CTAS=2
run ll128 NCCL_PROTO=LL128
run nonvls NCCL_NVLS_ENABLE=0
run nobucket X=1
timeout 200 python bench.py --steps 40 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_n2c_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_n2c_n1.json').read().strip().splitlines()[-1]); print('n1', round(d['ms_per_step'],4), '%.4e' % d['value'])"
