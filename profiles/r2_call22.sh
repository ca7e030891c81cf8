timeout 300 python profiles/dbg_dense_uf16.py > gpurun_out/r2_dbg_dense.log 2>&1; tail -80 gpurun_out/r2_dbg_dense.log
