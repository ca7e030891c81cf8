set -x
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_gpl --launch-skip 3 -c 1 -o gpurun_out/r2_prof16_gpl -f python profiles/subm_microbench.py > gpurun_out/r2_ncu16.log 2>&1; tail -2 gpurun_out/r2_ncu16.log
