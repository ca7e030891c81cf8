# Round-end validation on one B200: GPU test suite, smoke(), the bench line (with roofline + CPU baseline) and the
# reference arm.  (The ncu launch list / timeline of the same build: see README.md section 4.)
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cat gpurun_out/bench_r1_final.json
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_r1_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1_ref.json
timeout 200 python profiles/step_timeline.py > gpurun_out/timeline_r1_final.txt 2>&1
