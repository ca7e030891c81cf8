set -x
timeout 1700 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider 2>&1 | tail -120 > gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke1.log 2>&1; tail -5 gpurun_out/r2_smoke1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 1500 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 > gpurun_out/r2_bench1_c1.json 2> gpurun_out/r2_bench1_c1.err; tail -c 800 gpurun_out/r2_bench1_c1.json; tail -5 gpurun_out/r2_bench1_c1.err
