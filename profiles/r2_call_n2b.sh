set -x
python - <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_trainer as t
open("/tmp/dp_worker.py", "w").write(t._DP_WORKER % dict(root=t.ROOT))
PY
timeout -s INT 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 /tmp/dp_worker.py > gpurun_out/r2_n2_worker.log 2>&1; echo rc=$?; tail -30 gpurun_out/r2_n2_worker.log
