set -x
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rA 2>&1 | grep -E "dense uf16|passed|failed|FAILED|Error" | tail -12 > gpurun_out/r2_t24.log; cat gpurun_out/r2_t24.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench24_c2.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench24_c2.json
