set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_trainer.py -q -m gpu -p no:cacheprovider -k "two_rank" 2>&1 | tail -5 > gpurun_out/r2_n2_test.log; tail -5 gpurun_out/r2_n2_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; cut -c1-300 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err; cut -c1-300 gpurun_out/r2_bench_n2_ref.json; tail -2 gpurun_out/r2_bench_n2_ref.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_n1_samebox.json
