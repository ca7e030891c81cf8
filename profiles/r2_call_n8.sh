set -x
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_n8.json; wc -l gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --config c5 --steps 4 --warmup 3 --no-roofline > gpurun_out/r2_bench_n8_c5.json 2> gpurun_out/r2_bench_n8_c5.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_n8_c5.json; tail -3 gpurun_out/r2_bench_n8_c5.err | cut -c1-200
