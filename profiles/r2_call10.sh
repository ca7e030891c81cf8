set -x
timeout 300 python profiles/subm_microbench.py > gpurun_out/r2_micro10.json 2> gpurun_out/r2_micro10.err; cut -c1-400 gpurun_out/r2_micro10.json; tail -3 gpurun_out/r2_micro10.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_gpl --launch-skip 3 -c 1 -o gpurun_out/r2_prof10_gpl -f python profiles/subm_microbench.py > gpurun_out/r2_ncu10.log 2>&1; tail -3 gpurun_out/r2_ncu10.log
ls -la gpurun_out/
