set -x
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=10 2>&1 | tail -15 > gpurun_out/r2_t9.log; tail -6 gpurun_out/r2_t9.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke9.log 2>&1; tail -4 gpurun_out/r2_smoke9.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench9_c2.json 2> gpurun_out/r2_bench9_c2.err; cut -c1-260 gpurun_out/r2_bench9_c2.json; tail -3 gpurun_out/r2_bench9_c2.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench9_ref.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench9_ref.json
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --no-roofline > gpurun_out/r2_bench9_c1.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench9_c1.json
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench9_c3.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench9_c3.json
timeout 300 python bench.py --input-format compact --steps 20 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/r2_bench9_c2_compact.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench9_c2_compact.json
timeout 600 python profiles/dense_cudnn_bar.py --bs 4 > gpurun_out/r2_dense_cudnn_bs4.json 2> gpurun_out/r2_dense_cudnn_bs4.err || timeout 400 python profiles/dense_cudnn_bar.py --bs 2 > gpurun_out/r2_dense_cudnn_bs2.json 2> gpurun_out/r2_dense_cudnn_bs2.err
cat gpurun_out/r2_dense_cudnn_bs*.json; tail -2 gpurun_out/r2_dense_cudnn_bs*.err
timeout 400 python profiles/dense_cudnn_bar.py --bs 4 --no-tf32 > gpurun_out/r2_dense_cudnn_bs4_fp32.json 2>/dev/null; cat gpurun_out/r2_dense_cudnn_bs4_fp32.json
timeout 600 python profiles/dense_step.py --bs 4 --steps 2 > gpurun_out/r2_dense_ours_bs4.log 2>&1; tail -3 gpurun_out/r2_dense_ours_bs4.log
