set -x
nvidia-smi -L | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_n4.json; wc -l gpurun_out/r2_bench_n4.json; tail -3 gpurun_out/r2_bench_n4.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --config c5 --steps 4 --warmup 3 --no-roofline > gpurun_out/r2_bench_n4_c5.json 2> gpurun_out/r2_bench_n4_c5.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_n4_c5.json; tail -3 gpurun_out/r2_bench_n4_c5.err
