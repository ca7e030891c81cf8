set -x
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=10 2>&1 | tail -15 > gpurun_out/r2_t17.log; tail -6 gpurun_out/r2_t17.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke17.log 2>&1; tail -3 gpurun_out/r2_smoke17.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench17_c2.json 2> gpurun_out/r2_bench17_c2.err; cut -c1-260 gpurun_out/r2_bench17_c2.json; tail -3 gpurun_out/r2_bench17_c2.err
timeout 400 python bench.py --config c5 --steps 4 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_bench17_c5.json 2> gpurun_out/r2_bench17_c5.err; cut -c1-200 gpurun_out/r2_bench17_c5.json; tail -3 gpurun_out/r2_bench17_c5.err
