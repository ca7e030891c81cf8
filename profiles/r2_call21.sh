set -x
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu -p no:cacheprovider -rA -k uf16 2>&1 | grep -E "dense uf16|passed|failed|FAILED|Error" | tail -12 > gpurun_out/r2_t21.log; cat gpurun_out/r2_t21.log
