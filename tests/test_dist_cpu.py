"""world_size-2 gloo tests (CPU) of the data-parallel host logic: event sharding as the
reference's IO does it, and the one-flat-buffer SUM all-reduce reproducing the reference's
DataParallel gradient semantics (sum over replicas, BatchNorm statistics per replica)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from conftest import small_flags
    from helpers import event_blob
    from oracle.model import OracleUResNet, segmentation_loss
    from uresnet_pytorch_b200.trainer import allreduce_sum_, shard_events

    flags = small_flags(SPATIAL_SIZE=16, URESNET_FILTERS=4, URESNET_NUM_STRIDES=2)
    torch.manual_seed(0)
    net = OracleUResNet(flags).train()
    for t in list(net.parameters()) + list(net.buffers()):
        dist.broadcast(t.data, 0)
    # 4 events per step, 2 per rank, contiguous blocks
    data, label = event_blob(4, 16, seed=21)
    mine = shard_events(4, world, rank)
    sel = np.isin(data[:, -2].astype(int), mine)
    d, l = torch.from_numpy(data[sel]), torch.from_numpy(label[sel])
    loss, _ = segmentation_loss(net(d)[0], d, l)
    loss.backward()
    local = [p.grad.clone() for p in net.parameters()]
    grads = [p.grad for p in net.parameters()]
    allreduce_sum_(grads)
    torch.save({"local": local, "summed": [g.clone() for g in grads], "events": mine},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sum_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert r0["events"] == [0, 1] and r1["events"] == [2, 3]
    for a, b, s0, s1 in zip(r0["local"], r1["local"], r0["summed"], r1["summed"]):
        assert torch.equal(s0, s1)                       # every rank holds the same reduced gradient
        assert torch.allclose(s0, a + b, rtol=0, atol=0)  # SUM, not mean
    assert any((a - b).abs().max() > 0 for a, b in zip(r0["local"], r1["local"]))  # shards really differ


def test_shard_events_matches_reference_assignment():
    from uresnet_pytorch_b200.trainer import shard_events
    # new_id = data_id // batch_per_gpu  (iotools_sparse.py:140)
    for world, n in [(2, 8), (4, 8), (8, 64)]:
        per = n // world
        for r in range(world):
            assert shard_events(n, world, r) == [i for i in range(n) if i // per == r]
    with pytest.raises(ValueError):
        shard_events(7, 2, 0)


def test_bench_pool_is_size_bucketed_and_rank_specific():
    """bench.make_pool: every rank draws its own event stream and orders its pool by voxel count,
    so the ranks of one step hold events of similar size; --no-bucket keeps the generation order."""
    import bench

    class A(object):
        events_per_gpu, pool, no_bucket = 1, 4, False

    saved = dict(bench.CONFIG)
    bench.CONFIG.update(ss=64, dim=3)
    try:
        p0, p1 = bench.make_pool(A, 0), bench.make_pool(A, 1)
        A.no_bucket = True
        q0 = bench.make_pool(A, 0)
    finally:
        bench.CONFIG.update(saved)
    n0, n1 = [len(d) for d, _ in p0], [len(d) for d, _ in p1]
    assert n0 == sorted(n0) and n1 == sorted(n1) and n0 != n1
    assert sorted(len(d) for d, _ in q0) == n0  # same events, different order
    assert all(d.shape[1] == 5 and l.shape == (len(d), 1) for d, l in p0)


def test_bench_keeps_stdout_to_one_json_line():
    """bench.py's contract is ONE JSON line on stdout.  Libraries write banners to file descriptor 1 behind Python's
    back (NCCL prints its version there when the first communicator is created -- it preceded the JSON line of the
    2-GPU run): bench._own_stdout() points descriptor 1 at stderr and keeps the real stdout for Python's print."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; bench._own_stdout(); "
            "os.write(1, b'NCCL version 0.0.0+banner\\n'); os.system('echo from-a-child-process'); "
            "print(json.dumps({'ok': 1}))") % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines() == ['{"ok": 1}'], r.stdout
    assert "banner" in r.stderr and "from-a-child-process" in r.stderr


def test_bench_reference_arm_line_shape():
    """`bench.py --impl reference`: one line, the GPU arm's metric / unit / config, impl = reference, cpu_baseline of
    kind "port", e2e with zero copy bytes (the tier's reading of the bench contract)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200SCN_REF_BUDGET_S="5", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--config", "c1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "voxels/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"] and d["value"] > 0
