set -x
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu -p no:cacheprovider -rA 2>&1 | grep -E "dense uf16|passed|failed|FAILED|Error" | tail -12 > gpurun_out/r2_t20.log; cat gpurun_out/r2_t20.log
timeout 600 python -m pytest tests/test_gpu_mma.py -q -m gpu -p no:cacheprovider -k "centre_first or grouped_pair_lists" 2>&1 | tail -5 > gpurun_out/r2_t20b.log; cat gpurun_out/r2_t20b.log
