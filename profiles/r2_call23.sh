timeout 300 python profiles/dbg_dense_layers.py > gpurun_out/r2_dbg_dense_layers.log 2>&1; tail -30 gpurun_out/r2_dbg_dense_layers.log
