set -x
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2_t32.log; tail -3 gpurun_out/r2_t32.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench32_c2.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench32_c2.json
