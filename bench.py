#!/usr/bin/env python
"""bench.py -- sparse U-ResNet train voxels/sec at 512^3 on N B200s (+ subm-conv HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle port on host cores

One "step" = one full training step (rulebook build, forward, loss, backward, gradient
all-reduce, Adam) on one batch of synthetic LArTPC-like events per GPU -- BASELINE.json
configs[1]: uresnet_sparse 3D, ss=512, uns=5, uf=16, nc=5, bs=1 (README command of the
reference).  Weak scaling: every rank processes `--events-per-gpu` events per step.

Timing: W >= 3 warm-up steps, then K steps each bracketed by CUDA events on the launching
stream with a 256 MB L2 flush between steps (outside the events); barrier +
torch.cuda.synchronize() on both sides; MAX over ranks.  `value` = steps on inputs already
resident in HBM; `e2e` = the same metric through Trainer.train_step(host blob) with the
pinned H2D copy of the float64 blob and the D2H read of the loss inside the timed region.

`roofline` (rank 0, before the process group exists): the level-0 SubmanifoldConvolution 16->16 forward over 128
concatenated events (working set > L2), CUDA events on the launching stream, L2 flushed -- algorithmic bytes /
launch time against the measured copy bandwidth; `roofline.rulebook` times the hash build, the neighbour table
and the pair-list build of the same workload; `roofline.step_frac` relates the WHOLE step's compulsory bytes to
`ms_per_step`.  `cpu_baseline` (N = 1 only): the reference arm in a clean child process.  stdout carries the one
JSON line and nothing else (`_own_stdout`); at N > 1 the processes leave without tearing NCCL down.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sparse U-ResNet train voxels/sec at 512^3"
CONFIG = dict(ss=512, uns=5, uf=16, nc=5, dim=3)
WORKLOAD = "uresnet_sparse 3D train ss=512 uns=5 uf=16 nc=5 bs=%d per GPU (BASELINE configs[1]), Adam, fp32"
TRAIN = True
# other BASELINE.json configs, selectable with --config for parity / scaling runs (never the default line)
EXTRA_CONFIGS = {
    "c1": (dict(ss=192, uns=5, uf=16, nc=5, dim=3), 1,
           "uresnet_sparse 3D inference ss=192 uns=5 uf=16 nc=5 bs=%d per GPU (BASELINE configs[0]), fp32, "
           "softmax + argmax + per-event metrics out"),
    "c3": (dict(ss=512, uns=5, uf=16, nc=5, dim=2), 16,
           "uresnet_sparse 2D train ss=512 uns=5 uf=16 nc=5 bs=%d per GPU (BASELINE configs[2]), Adam, fp32"),
    "c5": (dict(ss=768, uns=6, uf=32, nc=5, dim=3), 8,
           "uresnet_sparse 3D train ss=768 uns=6 uf=32 nc=5 bs=%d per GPU (BASELINE configs[4]), Adam, fp32"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events-per-gpu", type=int, default=1)
    ap.add_argument("--pool", type=int, default=10, help="distinct synthetic steps cycled through (30 default steps = 3 passes)")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c5"],
                    help="c2 = BASELINE configs[1] (the metric's config, default); c1 / c3 / c5 = configs[0] / [2] / [4]")
    ap.add_argument("--no-bucket", action="store_true", help="keep the pool in generation order (no size bucketing)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="keep wgrad kernels on the main stream")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--roofline-events", type=int, default=128)
    ap.add_argument("--input-format", default="blob64", choices=["blob64", "compact"],
                    help="blob64: the reference's wire format (float64 [N, D+2]); compact: int32 coords + float32 values")
    ap.add_argument("--math", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="fp32: FFMA kernels; tf32x3: tensor-core 3-term split, fp32-equivalent (parity-tested at the "
                         "fp32 tolerances); tf32: single-pass tensor core (reduced precision)")
    args = ap.parse_args()
    if args.config != "c2":
        global WORKLOAD, METRIC, TRAIN
        cfg, bs, WORKLOAD = EXTRA_CONFIGS[args.config]
        CONFIG.update(cfg)
        TRAIN = args.config != "c1"
        METRIC = "sparse U-ResNet %s voxels/sec at %d^%d" % ("train" if TRAIN else "inference", cfg["ss"], cfg["dim"])
        if args.events_per_gpu == 1:
            args.events_per_gpu = bs
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_pool(args, rank):
    """`--pool` distinct synthetic steps per rank (rank r draws from its own stream).  Unless
    --no-bucket is given the pool is ordered by voxel count on every rank -- size-bucketed sampling: the ranks
    of one step then hold events of similar size, so the per-step gradient all-reduce waits less for the
    largest event (same events, same total work; at N=1 only the order changes)."""
    from uresnet_pytorch_b200 import synthetic
    stream = synthetic.EventStream(args.events_per_gpu, CONFIG["ss"], CONFIG["dim"], rank=rank)
    pool = [stream.next() for _ in range(args.pool)]
    if not getattr(args, "no_bucket", False):
        pool.sort(key=lambda dl: len(dl[0]))
    return pool


# ------------------------------------------------------------------------------------------------
def cpu_oracle_step_time(n_steps, events_per_step, threads):
    """The oracle port (oracle/, torch-CPU restatement of scn's per-offset gather/mm/scatter)
    timed on host cores: full train steps (or forward passes, for the inference config) of the same
    workload.  Baseline only."""
    import torch
    from oracle.model import OracleUResNet, segmentation_loss
    from uresnet_pytorch_b200 import synthetic
    from uresnet_pytorch_b200.model import Flags
    torch.set_num_threads(threads)
    flags = Flags(DATA_DIM=CONFIG["dim"], SPATIAL_SIZE=CONFIG["ss"], URESNET_FILTERS=CONFIG["uf"],
                  URESNET_NUM_STRIDES=CONFIG["uns"], NUM_CLASS=CONFIG["nc"])
    torch.manual_seed(0)
    net = OracleUResNet(flags).train(TRAIN)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    stream = synthetic.EventStream(events_per_step, CONFIG["ss"], CONFIG["dim"])
    times, vox = [], []
    for i in range(n_steps + 1):
        data, label = stream.next()
        d, l = torch.from_numpy(data), torch.from_numpy(label)
        t0 = time.perf_counter()
        if TRAIN:
            out = net(d)[0]
            loss, _ = segmentation_loss(out, d, l)
            opt.zero_grad()
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():   # R/uresnet/trainval.py:92-130 with TRAIN=False: logits, softmax, loss / accuracy
                out = net(d)[0]
                torch.softmax(out, 1)
                segmentation_loss(out, d, l)
        dt = time.perf_counter() - t0
        if i > 0:  # first step warms the allocator / thread pool
            times.append(dt)
            vox.append(len(data))
    return times, vox


def ref_events_per_step(args):
    """Events one reference-arm step processes: the GLOBAL batch of the GPU arm (events per GPU x GPUs), bounded
    so that K steps stay within minutes on host cores (a c2 event costs ~0.7 s, a c5 event ~10 s)."""
    want = args.events_per_gpu * max(args.gpus, 1)
    bound = {"c1": 16, "c2": 8, "c3": 16, "c5": 1}[args.config]
    return min(want, bound), want


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    ev, want = ref_events_per_step(args)
    budget = float(os.environ.get("B200SCN_REF_BUDGET_S", "150"))
    # bounded sample: stop early (and say so) rather than run for tens of minutes on a slow host
    t_start = time.perf_counter()
    times, vox = [], []
    import torch  # noqa: F401
    chunk = 1
    done = 0
    while done < steps:
        t, v = cpu_oracle_step_time(min(chunk, steps - done), ev, threads)
        times += t
        vox += v
        done += len(t)
        if time.perf_counter() - t_start > budget:
            break
        chunk = max(1, min(8, steps - done))
    total_t, total_v = sum(times), sum(vox)
    value = total_v / total_t
    what = "train steps" if TRAIN else "forward passes"
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": len(times), "warmup": 1,
        "ms_per_step": 1e3 * total_t / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.events_per_gpu, "events_per_gpu": args.events_per_gpu},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": "%d full %s (1 untimed per chunk) of the CPU oracle port, %d event(s) per step "
                                   "(GPU arm: %d per step over %d GPU(s)); scn is not installable here, so the port "
                                   "stands in for the reference's CPU path; ONE host process on all cores -- its "
                                   "voxels/s does not grow with --gpus" % (len(times), what, ev, want, args.gpus)},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def subm_roofline(args, math_mode):
    """subm-conv HBM GB/s on a batch large enough to exceed L2: level-0
    SubmanifoldConvolution 16->16 over `--roofline-events` concatenated 512^3 events, timed with
    CUDA events on the launching stream, L2 flushed between launches."""
    import torch
    from uresnet_pytorch_b200 import _lib, synthetic
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata
    dev = torch.device("cuda")
    data, _, _ = synthetic.make_batch(args.roofline_events, 512, 3, seed=4321)
    coords = torch.from_numpy(data[:, :4]).to(dev)
    md = Metadata(3)
    size = (512,) * 3
    lv = md.set_input(coords, size)
    nbr = md.subm_table(size, 3)
    n, C, K = lv.n, 16, 27
    torch.manual_seed(0)
    x = torch.randn(n, C, device=dev)
    W = torch.randn(K, C, C, device=dev) * 0.05
    out = torch.empty(n, C, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    pairs = int((nbr[:, :n] >= 0).sum().item())

    use_gpl = scn.config.use_gpl and math_mode != scn.MATH_FP32 and not os.environ.get("B200SCN_NO_GPL")
    gpl = md.subm_gpl(size, 3) if use_gpl else None
    if use_gpl:  # the list layout SubmanifoldConvolution.forward would pick for this shape (functional._conv_gpl)
        want = _lib.load().b200scn_conv_gpl_layout(gpl["tile_rows"], int(gpl["tile_end"] is None), n, C, C, K, 1, math_mode)
        if want:
            gpl = md.subm_gpl(size, 3, want)

    def launch():
        if use_gpl:  # what SubmanifoldConvolution.forward runs in the tensor-core math modes
            _lib.call("b200scn_conv_gpl", x.data_ptr(), C, W.data_ptr(), None, C * C, C, 1, 0, gpl["tile_rows"],
                      gpl["layout"], gpl["tile_off"].data_ptr(), None, gpl["g_rec"].data_ptr(), K, n, None, 0, C, C,
                      out.data_ptr(), C, 0, None, 0, math_mode, st.cuda_stream)
        else:
            _lib.call("b200scn_conv_table", x.data_ptr(), C, W.data_ptr(), C * C, C, 1, 0, nbr.data_ptr(),
                      nbr.stride(0), K, n, None, C, C, out.data_ptr(), C, 0, math_mode, st.cuda_stream)

    for _ in range(3):
        launch()
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        launch()
        e1.record(st)
        e1.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    t = statistics.mean(times)
    # algorithmic bytes: s*(N*Cin + N*Cout) + 4*K*N + s*K*Cin*Cout
    alg = 4 * (n * C + n * C) + 4 * K * n + 4 * K * C * C
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram__bytes_read+write per launch, from ncu --set full
    if os.path.exists(tpath) and args.roofline_events == 128:
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("subm16_fwd_bytes_per_launch"), tj.get("source")
    # the rulebook work behind that launch, same workload (north star, job 1): active-site hash table + row order,
    # 3^3 neighbour table, grouped pair list -- each timed as a whole call (CUDA events, best of 3 fresh builds, L2
    # flushed; the input build includes its one 8-byte read-back), against its algorithmic bytes
    def timed(fn):
        best = None
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            r = fn()
            e1.record(st)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best, r

    peak, src = peaks()
    rb = {}
    try:
        t_in, _ = timed(lambda: Metadata(3).set_input(coords, size))
        md2 = Metadata(3)
        md2.set_input(coords, size)
        t_tab, _ = timed(lambda: (md2.levels[size].subm.clear(), md2.subm_table(size, 3)))
        lay = gpl["layout"] if use_gpl else 0
        t_gpl, _ = timed(lambda: (md2.levels[size].gpl.clear(), md2.subm_gpl(size, 3, lay)))
        n_in = coords.shape[0]
        for name, ms, nbytes in (("input_build (hash insert, first-occurrence row order, duplicate CSR)", t_in,
                                  8 * 4 * n_in + 16 * 2 * n_in + 8 * n + 12 * n_in),
                                 ("subm_table (3^3 neighbour table)", t_tab, 8 * n + 4 * K * n),
                                 ("gpl_build (count + scan + fill, layout %d)" % lay, t_gpl,
                                  2 * 4 * K * n + (gpl["g_rec"].numel() * 4 if use_gpl else 0))):
            rb[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB/s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
        del md2
    except Exception as e:  # never lose the conv number over the side measurement
        rb = {"error": repr(e)}
    del flush
    torch.cuda.empty_cache()
    return {"bound": "hbm", "achieved": alg / t / 1e9, "rulebook": rb, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": src, "kernel": ("b200scn_conv_gpl" if use_gpl else "b200scn_conv_table") + " (subm 3^3, 16->16, fwd)",
            "workload": "%d events of 512^3 concatenated: %d active sites, %d rule pairs (working set > L2)"
                        % (args.roofline_events, n, pairs),
            "algorithmic_bytes_per_launch": alg, "launch_ms": t * 1e3, "gflops": 2.0 * pairs * C * C / t / 1e9,
            "pair_list_layout": gpl["layout"] if use_gpl else None}


def cpu_baseline_subprocess(args):
    """cpu_baseline leg (N = 1 only): the reference arm in a CLEAN child process -- no CUDA context, no NCCL,
    no torchrun environment (OMP_NUM_THREADS=1 and friends), so host threads behave as in a plain CPU run."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "MASTER_ADDR",
                        "MASTER_PORT", "OMP_NUM_THREADS", "MKL_NUM_THREADS", "TORCHELASTIC_RUN_ID")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["B200SCN_REF_BUDGET_S"] = "25"
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", "1", "--steps", "3",
           "--config", args.config, "--events-per-gpu", str(args.events_per_gpu)]
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["sample"] += "; %.2f s/step" % (line["ms_per_step"] / 1e3)
        return cb
    except Exception as e:  # the headline line must survive a broken baseline leg
        return {"value": None, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}


def run_b200(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference)")
    from uresnet_pytorch_b200 import _lib
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.model import Flags
    from uresnet_pytorch_b200.trainer import Trainer, init_distributed

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    math_mode = {"fp32": scn.MATH_FP32, "tf32x3": scn.MATH_TF32X3, "tf32": scn.MATH_TF32}[args.math]

    # Rank-0-only legs run BEFORE the process group exists: the other ranks then wait in the store rendezvous of
    # init_process_group (no NCCL collective pending, no watchdog), never in a barrier behind host work.
    roof = cpu = None
    if rank == 0:
        if not args.no_cpu_baseline and world_env == 1:
            cpu = cpu_baseline_subprocess(args)
        if not args.no_roofline:
            try:
                roof = subm_roofline(args, math_mode)
            except Exception as e:  # keep the headline line even if the microbench cannot run
                roof = {"bound": "hbm", "achieved": None, "peak": peaks()[0], "unit": "GB/s", "frac": None,
                        "traffic": None, "error": repr(e)}

    rank, local, world = init_distributed()
    dist = torch.distributed if world > 1 else None
    dev = torch.device("cuda", local)
    flags = Flags(DATA_DIM=CONFIG["dim"], SPATIAL_SIZE=CONFIG["ss"], URESNET_FILTERS=CONFIG["uf"],
                  URESNET_NUM_STRIDES=CONFIG["uns"], NUM_CLASS=CONFIG["nc"])
    pool = make_pool(args, rank)
    cap = max(len(d) for d, _ in pool)
    cap = ((cap + cap // 8 + 4095) // 4096) * 4096 * int(os.environ.get("B200SCN_BENCH_CAP_MULT", "1"))
    graph = not args.no_graph
    trainer = Trainer(flags, lr=1e-3, graph=graph, capacity=cap if graph else None, train=TRAIN,
                      events_per_step=args.events_per_gpu, math_mode=math_mode, seed=0,
                      overlap_wgrad=not args.no_overlap, input_format=args.input_format)

    def host_input(d):
        if args.input_format == "compact":
            D = CONFIG["dim"]
            return (np.ascontiguousarray(d[:, :D + 1], dtype=np.int32), np.ascontiguousarray(d[:, D + 1], dtype=np.float32))
        return d

    hpool = [(host_input(d), l) for d, l in pool]

    def public_step(i):
        """The call a user makes, synchronous form."""
        d, l = hpool[i % len(hpool)]
        if TRAIN:
            return trainer.train_step(d, l)
        return trainer.predict_step(d, l)

    launches0 = _lib.load().b200scn_launch_count()
    r = public_step(0)  # builds / captures the graph
    launches_capture = _lib.load().b200scn_launch_count() - launches0
    if TRAIN:
        assert np.isfinite(r["loss"]), r

    # device-resident copies of the pool (inputs in HBM before the timed region starts)
    dpool = []
    for (d, l), (raw, _) in zip(hpool, pool):
        n = len(raw)
        if args.input_format == "compact":
            dd = (torch.from_numpy(d[0]).to(dev), torch.from_numpy(d[1]).reshape(-1, 1).to(dev))
        else:
            dd = torch.from_numpy(d).to(dev)
        ll = torch.from_numpy(l).to(dev)
        dpool.append((dd, ll, torch.tensor([n], dtype=torch.int32, device=dev), n))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def resident_step(i):
        dd, ll, nn, n = dpool[i % len(dpool)]
        if graph:
            if args.input_format == "compact":
                trainer.d_coords[:n].copy_(dd[0])
                trainer.d_value[:n].copy_(dd[1])
            else:
                trainer.d_data[:n].copy_(dd)
            trainer.d_label[:n].copy_(ll)
            trainer.d_n.copy_(nn)
            if TRAIN:
                trainer.step_resident()
            else:
                trainer.predict_resident()
        elif TRAIN:
            trainer.step_resident(dd, ll)
        else:
            trainer._predict_dev(dd, ll)
        return n

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # per-replay launch count (graph: what one replay contains; eager: measured per step)
    l0 = _lib.load().b200scn_launch_count()
    for i in range(max(args.warmup, 3)):
        resident_step(i)
    torch.cuda.synchronize()
    per_step_eager = (_lib.load().b200scn_launch_count() - l0) / max(args.warmup, 3)
    per_step_launches = int(round(launches_capture / 4.0)) if graph else int(round(per_step_eager))
    # (graph mode: the first call ran 3 eager warm-up steps + 1 captured step => capture count / 4)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev, vox = [], 0
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        vox += resident_step(i)
        e1.record(st)
        ev.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    dev_s = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public API: host arrays -> pinned slot -> H2D -> step -> D2H of the result, every
    # step.  Training uses the pipelined pair submit()/result() (step i+1 is staged and uploaded while step i
    # runs on the GPU; every copy still lies inside the timed region); inference uses predict_step().
    for i in range(3):
        public_step(i)
    barrier()
    t0 = time.perf_counter()
    vox_e2e, h2d, d2h = 0, 0, 0

    def nbytes(d, l):
        return (sum(a.nbytes for a in d) if isinstance(d, tuple) else d.nbytes) + l.nbytes + 4

    if TRAIN:
        prev = None
        for i in range(args.steps):
            d, l = hpool[i % len(hpool)]
            t = trainer.submit(d, l)
            if prev is not None:
                trainer.result(prev)
            prev = t
            vox_e2e += len(l)
            h2d += nbytes(d, l)
            d2h += 24
        trainer.result(prev)
    else:
        for i in range(args.steps):
            d, l = hpool[i % len(hpool)]
            res = trainer.predict_step(d, l)
            vox_e2e += len(l)
            h2d += nbytes(d, l)
            d2h += res["softmax"].nbytes + res["pred"].nbytes + 8 * args.events_per_gpu * (4 + 5 + 50)
    barrier()
    e2e_s = time.perf_counter() - t0

    tt = torch.tensor([dev_s, e2e_s, wall], dtype=torch.float64, device=dev)
    vv = torch.tensor([vox, vox_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(vv, op=dist.ReduceOp.SUM)
    dev_s, e2e_s, wall = tt.tolist()
    vox_all, vox_e2e_all = vv.tolist()
    if dist is not None:
        dist.barrier()
    if rank != 0:
        return

    ms_step = 1e3 * dev_s / args.steps
    if roof is None:
        roof = {"bound": "hbm", "achieved": None, "peak": peaks()[0], "unit": "GB/s", "frac": None, "traffic": None,
                "note": "microbenchmark skipped (--no-roofline)"}
    # step-level fraction: compulsory bytes of the whole step (SURVEY.md 8d calculator: forward MB per event,
    # x3 for a training step) over the measured step time -- NOT the microbenchmark's number
    fwd_mb = {"c1": 106.0, "c2": 438.0, "c3": 136.0, "c5": 1710.0}[args.config]
    step_bytes = fwd_mb * 1e6 * (3 if TRAIN else 1) * args.events_per_gpu
    roof["step_algorithmic_bytes"] = step_bytes
    roof["step_frac"] = step_bytes / (ms_step * 1e-3) / 1e9 / roof["peak"]
    roof["step_note"] = ("step_frac = compulsory bytes of one whole step / ms_per_step / peak (the step is launch- and "
                         "latency-bound on L2-resident data); frac = the level-0 subm-conv kernel alone on a "
                         "working set larger than L2")

    line = {
        "metric": METRIC, "value": vox_all / dev_s, "unit": "voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32x3": "f32 (storage and accumulation fp32; products via 3-term tf32 split on mma.sync, fp32-equivalent: tests/test_gpu_bench_config.py)", "tf32": "tf32 (reduced precision)"}[args.math], "data": "synthetic",
        "config": {"workload": WORKLOAD % args.events_per_gpu, "events_per_gpu": args.events_per_gpu},
        "details": {"mean_voxels_per_step_per_gpu": vox_all / world / args.steps,
                    "cuda_graph": graph, "row_capacity": cap if graph else None, "input_format": args.input_format,
                    "sampling": "generation order" if args.no_bucket else "pool ordered by voxel count on every rank "
                                                                          "(size-bucketed steps)",
                    "l2": "256 MB buffer written between timed steps (L2 flush), outside the CUDA events",
                    "timing": "sum of per-step CUDA-event durations, max over ranks", "wall_s": wall},
        "e2e": {"value": vox_e2e_all / e2e_s, "unit": "voxels/s", "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": 1e3 * e2e_s / args.steps,
                "how": ("Trainer.submit(host blob, labels) / Trainer.result(): pinned staging + H2D on a copy stream + "
                        "step + D2H of (loss, acc, input-error counts), step i+1 staged while step i runs"
                        if TRAIN else
                        "Trainer.predict_step(host blob, labels): pinned staging + H2D + forward + D2H of softmax, "
                        "argmax and the per-event metric sums") + "; wall clock, barrier+sync both sides, max over ranks"},
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def _own_stdout():
    """The contract is ONE JSON line on stdout, but libraries write banners to file descriptor 1 behind Python's back
    (NCCL prints "NCCL version ..." there when the first communicator is created).  Python's sys.stdout keeps the
    real stdout; descriptor 1 itself is pointed at stderr for everything else in the process."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    args = parse()
    _own_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    # Every rank has passed the final barrier of run_b200 by now.  No NCCL teardown: ncclCommDestroy waits for every
    # CUDA graph that captured the communicator's kernels (tests/test_gpu_trainer.py met that hang), and nothing
    # is gained by an orderly teardown of a process that is about to exit.
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        multi = False
    if multi:
        os._exit(0)


if __name__ == "__main__":
    main()
