# Round-end validation on one B200: GPU test suite, smoke(), the bench line (with roofline + CPU baseline), the
# reference arm, the other BASELINE configs, then (only after those exited 0 without ncu) the ncu launch list of the
# bench command, one ncu --set full capture of the roofline kernel and the warm step timeline.
set -x
R=${R:-r2_final}
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/${R}_tests.log; tail -3 gpurun_out/${R}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -3 gpurun_out/${R}_smoke.log
timeout 600 python bench.py > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err; cut -c1-300 gpurun_out/${R}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${R}_bench_ref.json
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --no-roofline > gpurun_out/${R}_bench_c1.json 2>/dev/null; cut -c1-200 gpurun_out/${R}_bench_c1.json
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/${R}_bench_c3.json 2>/dev/null; cut -c1-200 gpurun_out/${R}_bench_c3.json
timeout 400 python bench.py --config c5 --steps 6 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/${R}_bench_c5.json 2>/dev/null; cut -c1-200 gpurun_out/${R}_bench_c5.json
timeout 300 python profiles/step_timeline.py > gpurun_out/${R}_timeline_c2.txt 2>&1; head -12 gpurun_out/${R}_timeline_c2.txt | tail -9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-roofline --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1; tail -2 gpurun_out/${R}_ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_gpl --launch-skip 3 -c 1 -o gpurun_out/${R}_prof_gpl -f python profiles/subm_microbench.py > gpurun_out/${R}_ncu_gpl.log 2>&1; tail -2 gpurun_out/${R}_ncu_gpl.log | cut -c1-200
