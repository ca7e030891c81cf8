set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cat gpurun_out/bench_r1_final.json
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_r1_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1_ref.json
timeout 200 python profiles/step_timeline.py > gpurun_out/timeline_r1_final.txt 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/plain8.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu19.log 2>&1; tail -1 gpurun_out/ncu19.log | cut -c1-200
