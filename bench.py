"""Headline benchmark: the full CycleGAN training step of the reference
(models/cycle_gan_model.py:138-160 — G_A/G_B ResNet-9 generators, D_A/D_B 70x70 PatchGANs, LSGAN + L1
cycle + identity losses, ImagePool 50, Adam, 4 D updates per G update) at batch 8, 256x256, on the B200
kernels of this repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cyclegan|g_infer|pix2pix|model5|metrics|segcycle] [--no-cuda-graph] [--no-batch-passes]

One JSON line on stdout (rank 0). See DESIGN.md "measurement" for the definition of every field.  The default
workload (the headline metric) shards the batch over N ranks under torchrun; `--workload metrics` shards the 697
image pairs; the other workloads run on one GPU.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "cyclegan_train_iters_per_s"
UNIT = "iters/s (batch-8 training steps at 256x256, summed over ranks)"
TFLOP_PER_SAMPLE = 2.1046   # 16.837 TFLOP per batch-8 step, counted on the reference modules
G_FWD_GFLOP = 99.10         # per image
DOMINANT = dict(n=8, c=256, hw=64, k=3)  # the 18 x 6 residual-block convolutions of a step


def make_opt(device, cuda_graph=False, batch_passes=True):
    return argparse.Namespace(cuda_graph=cuda_graph, batch_passes=batch_passes, input_nc=3, output_nc=3, ngf=64, ndf=64, netG='resnet_9blocks', netD='basic',
                              n_layers_D=3, norm='instance', no_dropout=True, init_type='normal', init_gain=0.02,
                              no_lsgan=False, pool_size=50, lr=2e-4, beta1=0.5, lambda_A=10.0, lambda_B=10.0,
                              lambda_identity=0.5, isTrain=True, device=device, direction='AtoB')


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="MEASURED_PEAKS.json")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def synthetic_batch(batch, size, seed):
    """size: side of a square image, or (height, width)."""
    h, w = size if isinstance(size, tuple) else (size, size)
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((batch, 3, h, w), generator=g) * 2 - 1,
            torch.rand((batch, 3, h, w), generator=g) * 2 - 1)


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle restatement of the reference step on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_builder(size, batch):
    from oracle import networks_oracle as O
    from cycle_depth_estimation_b200 import networks as N
    import contextlib
    import io
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        nets = [N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu']) for _ in range(2)]
        nets += [N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu']) for _ in range(2)]
    oracle = O.CycleGANStepOracle(*[n.state_dict() for n in nets])
    a, b = synthetic_batch(batch, size, 1234)
    return lambda: oracle.step(a, b)


def run_cpu_sample(steps, warmup, budget_s):
    """Times `steps` CPU steps of a bounded sample (batch 1; 256x256, or 128x128 when a 256 step would
    blow the budget). Returns (equivalent batch-8 256x256 iters/s, description, seconds per sample step)."""
    import random
    random.seed(1234)
    torch.set_num_threads(os.cpu_count() or 1)
    size = 256
    step = cpu_reference_step_builder(size, 1)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    used_warm = 1
    if first * (steps + max(warmup - 1, 0)) > budget_s:
        size = 128
        step = cpu_reference_step_builder(size, 1)
        used_warm = 0
    for _ in range(max(warmup - used_warm, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    scale = 8.0 * (256.0 / size) ** 2   # conv nets: work is linear in batch and in pixels
    return 1.0 / (dt * scale), "batch 1 at %dx%d per step, scaled by %.0fx to batch 8 at 256x256" % (size, size, scale), dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sample, dt = run_cpu_sample(args.steps, args.warmup, budget_s=200.0)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CycleGAN training step (reference restated on torch CPU), " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def time_dominant_kernel(peaks, iters=40):
    """CUDA-event timing of the dominant kernel (igemm_kernel on the 3x3 256->256 residual-block
    convolution at 64x64, batch 8: M=32768, N=256, K=2304) on torch's current stream."""
    from cycle_depth_estimation_b200 import ops
    d = DOMINANT
    n, c, hw, k = d["n"], d["c"], d["hw"], d["k"]
    x = torch.randn((n, hw + 2, hw + 2, c), device="cuda").to(torch.bfloat16)
    w = (torch.randn((c, c, k, k), device="cuda") * 0.02).contiguous()
    wp, rows_pad, kpad = ops.pack_conv_weight(w, True)
    y = ops.alloc_flat_output(n, hw, hw, hw + 2, c, "cuda")   # pitched like the engine's conv outputs
    stats = torch.zeros((n, c, 2), dtype=torch.float32, device="cuda")
    g = ops.geom(k, k)
    ov = ops.out_view_nhwc(y, c)
    for _ in range(5):
        ops.conv2d_fwd(g, x, wp, rows_pad, kpad, ov, None, ops.ACT_NONE, 0.0, stats)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv2d_fwd(g, x, wp, rows_pad, kpad, ov, None, ops.ACT_NONE, 0.0, stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * hw * hw * c * c * k * k
    achieved = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_burst"],
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the round-2 ncu --set full capture of this very
            # launch (profiles/r02_ncu_full_kernels.json, igemm_flat_kernel<0>, grid 136): 19.12 MB read, 0 written (the
            # 16.8 MB output stays in L2).  A profiler counter cannot be read inside an unprofiled run: the figure is the
            # committed capture's, not a live measurement; the algorithmic bytes are 17.8 MB input + 1.2 MB weights.
            "traffic": 19121664, "traffic_source": "profiles/r02_ncu_full_kernels.json (ncu --set full of this launch)",
            "kernel": "igemm_flat_kernel",
            "shape": "3x3 conv 256->256, 64x64, batch 8 (M=32768 N=256 K=2304), fused IN statistics",
            "us_per_launch": ms * 1e3, "peak_source": peaks["source"] + " (burst: kernel timed alone)"}


def time_tf32_kernel(iters=20):
    """The same 3x3 256->256 convolution through the TF32 variant (tcgen05 kind::tf32, fp32 NHWC in/out): reported
    in `config` next to the bf16 roofline. Peak: nominal dense TF32 (B200_PROFILING.md: 1.1 PFLOP/s) — there is no
    measured TF32 entry in MEASURED_PEAKS.json."""
    from cycle_depth_estimation_b200 import ops
    d = DOMINANT
    n, c, hw, k = d["n"], d["c"], d["hw"], d["k"]
    x = ops.round_tf32_(torch.randn((n, hw + 2, hw + 2, c), device="cuda"))
    w = (torch.randn((c, c, k, k), device="cuda") * 0.02).contiguous()
    wp, rows_pad, kpad = ops.pack_conv_weight_tf32(w, True)
    y = torch.empty((n, hw, hw, c), dtype=torch.float32, device="cuda")
    g = ops.geom(k, k)
    ov = ops.out_view_nhwc(y, c)
    for _ in range(3):
        ops.conv2d_fwd(g, x, wp, rows_pad, kpad, ov)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv2d_fwd(g, x, wp, rows_pad, kpad, ov)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * n * hw * hw * c * c * k * k / (ms * 1e-3) / 1e12
    return {"kernel": "igemm_flat_kernel (kind::tf32)", "us_per_launch": ms * 1e3, "tflops": tf,
            "frac_of_nominal_tf32_peak": tf / 1100.0, "peak_source": "nominal dense TF32 1.1 PFLOP/s (fallback)"}


def time_norm_backward(peaks, n=16, hw=64, c=256, pad=1, sets=3, rounds=4):
    """The second HBM-bound kernel family of the step: the InstanceNorm + ReLU backward of a residual-block layer
    (norm_bwd_tma_kernel reduce + apply, reflect fold of the padded gradient) at the batch the step runs it on (16).
    `sets` buffer sets (> 126 MB L2 together) are walked in turn inside one CUDA graph, so every launch streams from
    HBM and the Python launch rate is out of the measurement.  Algorithmic bytes: 2 reads + 1 write of the tensor."""
    from cycle_depth_estimation_b200 import ops
    bufs = []
    for _ in range(sets):
        y = ops.alloc_flat_output(n, hw, hw, hw + 2, c, "cuda")
        y.normal_()
        stats = torch.zeros((n, c, 2), device="cuda")
        ops.channel_stats(y, c, True, stats)
        dfull = torch.randn((n, hw + 2 * pad, hw + 2 * pad, c), device="cuda").to(torch.bfloat16)
        dyp = torch.zeros((n, hw + 4, hw + 4, c), dtype=torch.bfloat16, device="cuda")
        bst = torch.zeros((n, c, 2), device="cuda")
        desc = ops.norm_desc(ops.NORM_INSTANCE, ops.ACT_RELU, 0.0, 1e-5, c, pad, stats)
        bufs.append((desc, y, dyp[:, 2:2 + hw, 2:2 + hw, :], dfull[:, pad:pad + hw, pad:pad + hw, :], bst, stats))

    def once():
        for desc, y, dy, dinner, bst, _ in bufs:
            ops.norm_act_bwd(desc, y, dy, dinner, None, bst, None)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        once()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(rounds):
            once()
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 5
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * rounds * sets)
    t_bytes = n * hw * hw * c * 2
    return {"kernel": "norm_bwd_tma_kernel reduce + apply", "shape": "InstanceNorm + ReLU backward, %dx%dx%d, pad %d, batch %d"
            % (hw, hw, c, pad, n), "us_per_call": us, "algorithmic_bytes": 3 * t_bytes, "bytes_moved": 5 * t_bytes,
            "algorithmic_gbs": 3 * t_bytes / us * 1e-3, "moved_gbs": 5 * t_bytes / us * 1e-3,
            "frac_of_hbm_peak_algorithmic": 3 * t_bytes / us * 1e-3 / peaks["hbm"],
            "frac_of_hbm_peak_moved": 5 * t_bytes / us * 1e-3 / peaks["hbm"], "hbm_peak_gbs": peaks["hbm"],
            "peak_source": peaks["source"], "launches_per_call": 2,
            "how": "%d buffer sets of %.0f MB walked in turn inside one CUDA graph (no L2 reuse between calls)"
            % (sets, 3 * t_bytes / 1e6)}


def cudnn_same_box(batch, size, steps=3, warmup=2):
    """The bar SURVEY 2.3 / BASELINE.md 3 name: the SAME training step through stock PyTorch on this GPU — the
    oracle restatement of models/cycle_gan_model.py:80-160 around torch.nn.functional, i.e. cuDNN convolutions and
    ATen norm / loss / Adam kernels — timed with CUDA events in this run, (a) fp32 with TF32 allowed (PyTorch's
    default for convolutions on this hardware), (b) torch.autocast(bfloat16) with channels_last tensors.  The oracle is
    used here as a timed baseline only; nothing of it is on the B200 path."""
    import contextlib
    import io
    import random
    from oracle import networks_oracle as O
    from cycle_depth_estimation_b200 import networks as N
    out = {}
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        nets = [N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu']) for _ in range(2)]
        nets += [N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu']) for _ in range(2)]
    sds = [{k: v.cuda() for k, v in n.state_dict().items()} for n in nets]
    a, b = synthetic_batch(batch, size, 1234)
    a, b = a.cuda(), b.cuda()
    old_bench, old_tf32 = torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.benchmark = True
    try:
        for name in ("ms_fp32_tf32", "ms_bf16_autocast"):
            random.seed(1234)
            oracle = O.CycleGANStepOracle(*sds)
            xa, xb = a, b
            if name == "ms_bf16_autocast":
                xa = a.contiguous(memory_format=torch.channels_last)
                xb = b.contiguous(memory_format=torch.channels_last)
                for sd in (oracle.G_A, oracle.G_B, oracle.D_A, oracle.D_B):
                    for v in sd.values():
                        if v.dim() == 4:
                            v.data = v.data.contiguous(memory_format=torch.channels_last)
                ctx = torch.autocast("cuda", dtype=torch.bfloat16)
            else:
                torch.backends.cudnn.allow_tf32 = True
                ctx = contextlib.nullcontext()
            with ctx:
                for _ in range(warmup):
                    oracle.step(xa, xb)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    oracle.step(xa, xb)
                e1.record()
                torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / steps
            del oracle
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = old_bench, old_tf32
    out["how"] = ("oracle.CycleGANStepOracle (torch.nn.functional / cuDNN %s, torch %s) on cuda:0, same weights shape, "
                  "inputs and step definition, %d warm-up + %d timed steps, CUDA events; cudnn.benchmark on; each "
                  "oracle step also reads its losses back (8 floats)" % (torch.backends.cudnn.version(), torch.__version__,
                                                                       warmup, steps))
    return out


def time_g_inference(batch, size=256, iters=30):
    """BASELINE configs[0] through the TestModel mirror (models/test_model.py:33-46): eager launches and CUDA-graph
    replay of the forward pass, device-resident input. Returns ms per forward for both."""
    import contextlib
    import io
    from cycle_depth_estimation_b200.test_model import TestModel
    res = {}
    a, _ = synthetic_batch(batch, size, 1234)
    dev = {'A': a.cuda(), 'A_paths': None}
    for mode in ("eager", "graph"):
        opt = make_opt("cuda", mode == "graph")
        opt.isTrain = False
        torch.manual_seed(0)
        m = TestModel()
        with contextlib.redirect_stdout(io.StringIO()):
            m.initialize(opt)
        m.eval()
        for _ in range(4):
            m.set_input(dev)
            m.test()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m.set_input(dev)
            m.test()
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / iters
        del m
    return res


def run_secondary_lines(steps=5):
    """The other BASELINE configs as sub-runs of this bench (one process each, this GPU), so that their numbers are
    recorded with the headline line instead of being builder-only claims."""
    out = []
    for wl in ("g_infer", "pix2pix", "model5", "metrics", "segcycle"):
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", str(steps), "--warmup", "3",
               "--no-cpu-baseline"]
        try:
            proc = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            line = json.loads([l for l in proc.stdout.splitlines() if l.startswith("{")][-1])
            out.append({"workload": wl, "metric": line["metric"], "value": line["value"], "unit": line["unit"],
                        "ms_per_step": line["ms_per_step"], "e2e_value": line["e2e"]["value"],
                        "roofline_frac": line["roofline"]["frac"], "roofline_bound": line["roofline"]["bound"],
                        "gpu_launches": line["gpu_launches"]})
        except Exception as exc:  # noqa: BLE001
            out.append({"workload": wl, "error": "%s: %s" % (type(exc).__name__, exc)})
    return out


def b200_arm(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cycle_depth_estimation_b200 import _lib
    from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel
    import contextlib
    import io
    import random

    lib = _lib.lib()
    peaks = load_peaks()
    batch, size = args.batch, args.size
    img_h, img_w = args.size, (args.width or args.size)      # --width: the fork's 192 x 640 training shape (config 2b)
    if img_w != img_h:
        size = (img_h, img_w)
    px_scale = img_h * img_w / 65536.0                       # TFLOP_PER_SAMPLE is counted at 256 x 256
    if args.scaling == "strong":
        # north_star: "the batch sharded" — the GLOBAL batch stays at --batch, every rank takes batch / world samples
        if batch % world:
            raise SystemExit("--scaling strong needs --batch divisible by the number of ranks")
        batch //= world
    torch.manual_seed(0)            # identical initial weights on every rank
    random.seed(1234)               # identical ImagePool stream on every rank
    host_a, host_b = synthetic_batch(batch, size, 1234 + rank)
    host_a, host_b = host_a.pin_memory(), host_b.pin_memory()
    dev = {"img_source": host_a.cuda(), "img_target": host_b.cuda()}

    def build(use_graph):
        torch.manual_seed(0)
        random.seed(1234)
        m = CycleGANModel()
        with contextlib.redirect_stdout(io.StringIO()):
            m.initialize(make_opt("cuda", use_graph, not args.no_batch_passes))
        return m

    # The whole step is replayed as ONE CUDA graph (cycle_gan_model.py) — under data parallelism the NCCL
    # all-reduces / all-gathers are captured with it; --no-cuda-graph issues eager launches. A failed
    # single-process capture restarts the process without graphs.
    use_graph = not args.no_cuda_graph
    graph_note = "cuda graph replay of the whole step" if use_graph else "eager launches"
    model = build(use_graph)
    if use_graph:
        try:
            for _ in range(CycleGANModel.GRAPH_WARMUP_STEPS + 2):
                model.set_input(dev)
                model.optimize_parameters("train")
            torch.cuda.synchronize()
            if model._graph is None:
                raise RuntimeError("step was not captured")
        except Exception as exc:  # noqa: BLE001
            print("cuda graph capture failed (%s: %s); falling back to eager launches" % (type(exc).__name__, exc),
                  file=sys.stderr)
            # a failed capture leaves the CUDA context / RNG in capture state: restart the process without graphs
            sys.stderr.flush()
            if world > 1:
                raise
            os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-cuda-graph"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    for _ in range(args.warmup):
        model.set_input(dev)
        model.optimize_parameters("train")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.cdb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        model.set_input(dev)
        model.optimize_parameters("train")
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.cdb_launch_count() - launches0
    if use_graph:
        # replays do not pass through the library's host entry points: count the kernels of one captured step
        # (the launches the library recorded while the step was being captured) times the replays
        launches = int(getattr(model, "_graph_launches", 0)) * args.steps
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * 1e3 / ms_step * (batch / 8.0)   # batch-8 steps per second over all ranks

    # ---- end to end: pinned host inputs every step, losses read back every step
    host = {"img_source": host_a, "img_target": host_b}
    model.set_input(host)
    model.optimize_parameters("train")
    model.get_current_losses()
    barrier()
    e0.record()
    for _ in range(args.steps):
        model.set_input(host)
        model.optimize_parameters("train")
        losses = model.get_current_losses()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = {"value": world * 1e3 / ms_e2e * (batch / 8.0), "unit": UNIT,
           "h2d_bytes_per_step": int(host_a.numel() * 4 * 2), "d2h_bytes_per_step": 4 * len(losses),
           "ms_per_step": ms_e2e}
    abort = lib.cdb_device_abort_flag()

    def finish():
        """Data-parallel teardown. With a captured step (NCCL kernels inside the graph) alive, a further barrier /
        destroy_process_group was observed to block on this stack; every timed region above already ended with a
        barrier, so the ranks simply flush and leave (process exit releases the communicator)."""
        if world > 1:
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            if use_graph:
                os._exit(0)
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        finish()
        return
    # ---- generator inference (BASELINE configs[0]) and the dominant kernel, rank 0 only
    del model
    torch.cuda.empty_cache()
    g1, g8 = time_g_inference(1), time_g_inference(8)
    roof = time_dominant_kernel(peaks)
    tf32 = time_tf32_kernel()
    norm_bwd = time_norm_backward(peaks)
    step_tflops = TFLOP_PER_SAMPLE * px_scale * batch / (ms_step * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sample, _ = run_cpu_sample(steps=2, warmup=1, budget_s=40.0)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": "CycleGAN training step: G_A/G_B resnet_9blocks + D_A/D_B 70x70 PatchGAN, LSGAN + L1 "
                        "cycle/identity, ImagePool 50, Adam, 4 D updates per G update; batch %d per GPU at %dx%d "
                        "(BASELINE configs[1]%s)" % (batch, img_h, img_w, "" if px_scale == 1.0 else
                                                      "; other image shape, TFLOP scaled by the pixel count"),
            "per_gpu_batch": batch, "image": img_h if img_w == img_h else [img_h, img_w], "parallelism": "dp%d" % world,
            "launch_mode": graph_note,
            "batched_passes": not args.no_batch_passes,
            "l2": "working set per step (saved activations of 6 generator + 18 discriminator passes, > 5 GB) "
                  "exceeds the 126 MB L2; no explicit flush",
            "algorithmic_tflop_per_step": TFLOP_PER_SAMPLE * px_scale * batch,
            "step_tflops_per_gpu": step_tflops,
            "step_frac_of_sustained_bf16_peak": step_tflops / peaks["bf16_sustained"],
            "g_forward_img_per_s_batch1": 1e3 / g1["graph"],
            "g_forward_tflops_batch1": G_FWD_GFLOP / g1["graph"],
            "g_forward_ms": {"batch1_eager": g1["eager"], "batch1_graph_replay": g1["graph"],
                             "batch8_eager": g8["eager"], "batch8_graph_replay": g8["graph"]},
            "g_forward_img_per_s_batch8": 8e3 / g8["graph"],
            "g_forward_tflops_batch8": 8 * G_FWD_GFLOP / g8["graph"],
            "tf32_variant_r256_conv": tf32,
            "norm_backward_r256": norm_bwd,
            "device_abort_flag": abort,
        },
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_cudnn_baseline:
        try:
            cd = cudnn_same_box(batch, size)
            cd["speedup_vs_fp32_tf32"] = cd["ms_fp32_tf32"] / ms_step
            cd["speedup_vs_bf16_autocast"] = cd["ms_bf16_autocast"] / ms_step
        except Exception as exc:  # noqa: BLE001
            cd = {"error": "%s: %s" % (type(exc).__name__, exc)}
        line["config"]["cudnn_same_box"] = cd
    if world == 1 and not args.no_secondary:
        torch.cuda.empty_cache()
        line["secondary"] = run_secondary_lines()
    print(json.dumps(line), flush=True)
    finish()


# ---------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE configs[0], [2], [3], [4]): same JSON schema, selected with --workload
# ---------------------------------------------------------------------------------------------------
def _time_steps(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _model5_batch(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(s, generator=g) * 2 - 1
    seg_syn = torch.randint(0, 28, (b, 1, h, w), generator=g)
    seg_real = torch.randint(0, 28, (b, 1, h, w), generator=g)
    seg_real[torch.rand((b, 1, h, w), generator=g) < 0.02] = 255
    dls = r(b, 4, h, w)
    dls[dls > 0.9] = 1.0
    dls[dls < -0.9] = -1.0
    return {'img_real': r(b, 3, h, w), 'img_syn': r(b, 3, h, w), 'seg_l_real': seg_real, 'seg_l_syn': seg_syn,
            'dep_l_syn': r(b, 1, h, w), 'depth_l_s': dls}



def secondary_cpu_baseline(wl, batch):
    """Bounded CPU sample of the same workload on the host cores through the oracle restatements (kind "port":
    the reference is pure Python, oracle/ restates its arithmetic on torch / numpy). Returns the dict for the
    JSON line; the value is scaled to the unit of the GPU line."""
    import numpy as np
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    if wl == "metrics":
        from oracle import networks_oracle as O
        rng = np.random.default_rng(7)
        k = 12
        gt = rng.integers(0, 80, (k, 375, 1242), dtype=np.uint8)
        gt[rng.random((k, 375, 1242)) < 0.3] = 0
        pred = rng.integers(0, 256, (k, 375, 1242), dtype=np.uint8)
        t0 = time.perf_counter()
        O.eval_metric_arrays(list(gt), list(pred))
        dt = time.perf_counter() - t0
        return {"value": k / dt, "unit": "images/s", "cores": 1, "kind": "port",
                "sample": "%d of the 697 image pairs through the numpy restatement of my_eval.py (single thread, as the "
                          "reference)" % k}
    if wl == "g_infer":
        from oracle import networks_oracle as O
        from cycle_depth_estimation_b200 import networks as N
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            net = N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu'])
        sd = net.state_dict()
        x, _ = synthetic_batch(1, 256, 1234)
        with torch.no_grad():
            O.resnet_generator(sd, x, 9)
            t0 = time.perf_counter()
            for _ in range(3):
                O.resnet_generator(sd, x, 9)
            dt = (time.perf_counter() - t0) / 3
        return {"value": 1.0 / dt, "unit": "img/s", "cores": cores, "kind": "port",
                "sample": "3 forward passes of one 256x256 image (torch CPU fp32)"}
    if wl == "pix2pix":
        from oracle import networks_oracle as O
        from cycle_depth_estimation_b200 import networks as N
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            g = N.define_G(3, 3, 64, 'unet_256', 'batch', False, 'normal', 0.02, ['cpu'])
            d = N.define_D(6, 64, 'basic', 3, 'batch', True, 'normal', 0.02, ['cpu'])
        oracle = O.Pix2PixStepOracle(g.state_dict(), d.state_dict(), num_downs=8)
        a, b = synthetic_batch(2, 256, 1234)
        oracle.step(a, b)
        t0 = time.perf_counter()
        oracle.step(a, b)
        dt = time.perf_counter() - t0
        return {"value": 1.0 / (dt * batch / 2.0), "unit": "iters/s (scaled to batch %d)" % batch, "cores": cores,
                "kind": "port", "sample": "one batch-2 step at 256x256 (torch CPU fp32), scaled linearly to batch %d" % batch}
    if wl == "model5":
        from oracle import networks5_oracle as O5
        from cycle_depth_estimation_b200 import networks5_ds as N5
        mods = [N5.G_1(), N5.General_net(), N5.R_dep(), N5._Discriminator(512), N5._Discriminator(256),
                N5._Discriminator(128)]
        sds = [O5.synth_state_dict(m.state_dict(), i) for i, m in enumerate(mods)]
        for sd in sds[3:]:
            sd['model.1.weight'] = sd['model.10.weight']
        oracle = O5.SegDepthStepOracle(*sds)
        data = _model5_batch(1, 192, 640, 90)
        args5 = (data['img_syn'], data['img_real'], data['seg_l_syn'].squeeze(1), data['seg_l_real'].squeeze(1),
                 data['dep_l_syn'].squeeze(1), data['depth_l_s'])
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            oracle.step(*args5)
            dt = time.perf_counter() - t0
        return {"value": 1.0 / (dt * batch), "unit": "iters/s (scaled to batch %d)" % batch, "cores": cores,
                "kind": "port", "sample": "one batch-1 step at 192x640 (torch CPU fp32, no warm-up), scaled linearly to "
                                          "batch %d" % batch}
    if wl == "segcycle":
        from oracle import encoder_decoder_oracle as OE
        from cycle_depth_estimation_b200 import encoder_decoder as E
        from cycle_depth_estimation_b200 import networks as N
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            gs = [N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu']) for _ in range(2)]
            ds = [N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu']) for _ in range(2)]
        tn = [E._UNetEncoder(3), E._UNetEncoder(3), E._UNetDecoder(22), E._UNetDecoder(28)]
        oracle = OE.SegCycleStepOracle(*[m.state_dict() for m in gs + ds + tn])
        a, b2 = synthetic_batch(1, 256, 1234)
        la, lb = _seg_labels(1, 256, 22, 5), _seg_labels(1, 256, 28, 6)
        t0 = time.perf_counter()
        oracle.step(a, b2, la, lb)
        dt = time.perf_counter() - t0
        return {"value": 1.0 / (dt * batch), "unit": "iters/s (scaled to batch %d)" % batch, "cores": cores,
                "kind": "port", "sample": "one batch-1 step at 256x256 (torch CPU fp32, no warm-up), scaled linearly to "
                                          "batch %d" % batch}
    return None


def _seg_labels(n, size, classes, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, classes, (n, 1, size, size), generator=g)
    lab[torch.rand((n, 1, size, size), generator=g) < 0.02] = 255
    return lab


def secondary_arm(args):
    """pix2pix step (configs[2]), seg/depth step (configs[3]), generator inference (configs[0]), depth metrics
    (configs[4]) on one GPU. FLOP / byte figures: SURVEY 8(d)."""
    import contextlib
    import io
    import random
    import numpy as np
    from cycle_depth_estimation_b200 import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    lib = _lib.lib()
    peaks = load_peaks()
    torch.manual_seed(0)
    random.seed(1234)
    wl = args.workload
    mp_world, mp_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    mp_dist = None
    if mp_world > 1:
        if wl not in ("metrics", "pix2pix"):
            raise SystemExit("--workload %s runs on one GPU; the default CycleGAN workload, --workload pix2pix (BatchNorm "
                             "statistics all-reduced per layer) and --workload metrics shard over ranks" % wl)
        import torch.distributed as mp_dist
        mp_dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    launches0 = lib.cdb_launch_count()
    extra = {}
    if wl == "pix2pix":
        from cycle_depth_estimation_b200.pix2pix_model import Pix2PixModel
        b = args.batch if args.batch != 8 else 16
        opt = argparse.Namespace(input_nc=3, output_nc=3, ngf=64, ndf=64, netG='unet_256', netD='basic', n_layers_D=3,
                                 norm='batch', no_dropout=False, init_type='normal', init_gain=0.02, no_lsgan=True,
                                 pool_size=0, lr=2e-4, beta1=0.5, lambda_L1=100.0, isTrain=True, device='cuda',
                                 direction='AtoB')
        model = Pix2PixModel()
        with contextlib.redirect_stdout(io.StringIO()):
            model.initialize(opt)
        a, bb = synthetic_batch(b, 256, 1234)
        if mp_world > 1:
            # strong scaling: the global batch is sharded; every BatchNorm layer all-reduces its (sum, sum of squares)
            # forward and (sum dy, sum dy*xhat) backward, so the step equals the single-device one
            per = b // mp_world
            a, bb = a[mp_rank * per:(mp_rank + 1) * per].contiguous(), bb[mp_rank * per:(mp_rank + 1) * per].contiguous()
            extra["parallelism"] = "dp%d: global batch %d sharded, BatchNorm statistics all-reduced per layer" % (mp_world, b)
        host = {'A': a.pin_memory(), 'B': bb.pin_memory(), 'A_paths': None}
        dev = {'A': a.cuda(), 'B': bb.cuda(), 'A_paths': None}

        def step():
            model.set_input(dev)
            model.optimize_parameters()

        def step_e2e():
            model.set_input(host)
            model.optimize_parameters()
            return model.get_current_losses()
        ms = _time_steps(step, args.steps, args.warmup)
        ms_e2e = _time_steps(step_e2e, args.steps, 1)
        if mp_world > 1:
            t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
            mp_dist.all_reduce(t, op=mp_dist.ReduceOp.MAX)
            ms, ms_e2e = float(t[0]), float(t[1])
        tflop = 1.391 * b / 16.0
        metric, unit = "pix2pix_train_iters_per_s", "iters/s (batch-%d training steps at 256x256)" % b
        workload = ("pix2pix training step: UnetGenerator unet_256 (BatchNorm, dropout) + NLayerDiscriminator n_layers=3 on "
                    "cat(A,B), BCE GAN + 100 L1, Adam; batch %d at 256x256 (BASELINE configs[2])" % b)
        h2d, d2h = int(a.numel() * 4 * 2), 16
    elif wl == "model5":
        from cycle_depth_estimation_b200.model5 import Seg_Depth
        b = args.batch
        model = Seg_Depth()
        with contextlib.redirect_stdout(io.StringIO()):
            model.initialize(argparse.Namespace(lr=2e-4, beta1=0.5, pool_size=50, cuda_graph=not args.no_cuda_graph))
        extra["launch_mode"] = "eager launches" if args.no_cuda_graph else "cuda graph replay of the whole step"
        data = _model5_batch(b, 192, 640, 90)
        host = {k: v.pin_memory() for k, v in data.items()}
        dev = {k: v.cuda() for k, v in data.items()}

        def step():
            model.set_input(dev, 'train')
            model.optimize_parameters('train')

        def step_e2e():
            model.set_input(host, 'train')
            model.optimize_parameters('train')
            return model.get_current_losses()
        ms = _time_steps(step, args.steps, max(args.warmup, 5))
        ms_e2e = _time_steps(step_e2e, args.steps, 1)
        tflop = 3.251 * b
        metric, unit = "seg_depth_train_iters_per_s", "iters/s (batch-%d training steps at 192x640)" % b
        workload = ("new_multi model5 step: G_1 + General_net + R_dep + 3 feature discriminators, CE + L1 + BCEDep + "
                    "LSGAN, 8 optimizer updates; batch %d at 192x640 (BASELINE configs[3])" % b)
        h2d, d2h = int(sum(v.numel() * v.element_size() for v in data.values())), 32
    elif wl == "segcycle":
        from cycle_depth_estimation_b200.seg_cycle import SegCycle
        b = args.batch
        model = SegCycle()
        with contextlib.redirect_stdout(io.StringIO()):
            model.initialize(make_opt("cuda", not args.no_cuda_graph, not args.no_batch_passes))
        extra["launch_mode"] = "eager launches" if args.no_cuda_graph else "cuda graph replay of the whole step"
        a, bb = synthetic_batch(b, args.size, 1234)
        data = {'img_source': a, 'img_target': bb, 'lab_source': _seg_labels(b, args.size, 22, 5),
                'lab_target': _seg_labels(b, args.size, 28, 6)}
        host = {k: v.pin_memory() for k, v in data.items()}
        dev = {k: v.cuda() for k, v in data.items()}

        def step():
            model.set_input(dev)
            model.optimize_parameters('train')

        def step_e2e():
            model.set_input(host)
            model.optimize_parameters('train')
            return model.get_current_losses()
        ms = _time_steps(step, args.steps, max(args.warmup, 5))
        ms_e2e = _time_steps(step_e2e, args.steps, 1)
        # FlopCounterMode over the reference modules on the meta device: CycleGAN step with ONE D update 1.879
        # TFLOP/sample (15.034 / 8) + the four task-network passes fwd+bwd 0.923 TFLOP/sample
        tflop = (1.879 + 0.9235) * b * (args.size / 256.0) ** 2
        metric, unit = "segcycle_train_iters_per_s", "iters/s (batch-%d training steps at %dx%d)" % (b, args.size, args.size)
        workload = ("SegCycle step (models/seg_cycle.py, SURVEY 8(f) f3): CycleGAN resnet_9blocks + PatchGANs + two "
                    "U-Net task networks (encoder A/B, decoder 22/28 classes), LSGAN + L1 + 4 CrossEntropy, Adam, one D "
                    "update; batch %d at %dx%d" % (b, args.size, args.size))
        h2d, d2h = int(sum(v.numel() * v.element_size() for v in data.values())), 48
    elif wl == "g_infer":
        from cycle_depth_estimation_b200.test_model import TestModel
        b = args.batch if args.batch != 8 else 1
        opt = make_opt("cuda", not args.no_cuda_graph)
        opt.isTrain = False
        model = TestModel()
        with contextlib.redirect_stdout(io.StringIO()):
            model.initialize(opt)
        model.eval()
        extra["launch_mode"] = "eager launches" if args.no_cuda_graph else "cuda graph replay of the forward pass"
        a, _ = synthetic_batch(b, 256, 1234)
        host, dev = {'A': a.pin_memory(), 'A_paths': None}, {'A': a.cuda(), 'A_paths': None}

        def step():
            model.set_input(dev)
            model.test()

        def step_e2e():
            model.set_input(host)        # pinned host image -> device (static input buffer of the captured forward)
            model.test()
            return model.fake_B.cpu()    # generated image back to the host
        ms = _time_steps(step, args.steps, max(args.warmup, 4))
        ms_e2e = _time_steps(step_e2e, args.steps, 2)
        tflop = 0.0991 * b
        metric, unit = "resnet9_generator_img_per_s", "img/s (256x256 forward, batch %d)" % b
        workload = ("TestModel (models/test_model.py): ResnetGenerator resnet_9blocks ngf=64 InstanceNorm forward, batch %d, "
                    "256x256 (BASELINE configs[0])" % b)
        h2d, d2h = int(a.numel() * 4), int(a.numel() * 4)
        extra["images_per_step"] = b
    elif wl == "metrics":
        from cycle_depth_estimation_b200 import my_eval
        n_img, h, w = 697, 375, 1242
        rng = np.random.default_rng(2019)
        gt = rng.integers(0, 80, (n_img, h, w), dtype=np.uint8)
        gt[rng.random((n_img, h, w)) < 0.3] = 0
        pred = rng.integers(0, 256, (n_img, h, w), dtype=np.uint8)
        hg, hp = torch.from_numpy(gt).pin_memory(), torch.from_numpy(pred).pin_memory()
        # one process per GPU: rank r owns the images r, r + world, ... (strong scaling over the
        # fixed 697 pairs); the device-resident figure is max-over-ranks of the shard's kernel time, the end-to-end
        # figure goes through my_eval.eval_metric_arrays, which shards, all-gathers the 697 x 7 float32 rows and
        # reduces them in image order on every rank
        dg, dp = hg[mp_rank::mp_world].cuda(), hp[mp_rank::mp_world].cuda()
        from cycle_depth_estimation_b200 import ops
        ms = _time_steps(lambda: ops.depth_metrics(dg, dp), args.steps, args.warmup)
        ms_e2e = _time_steps(lambda: my_eval.eval_metric_arrays(hg, hp), args.steps, 1)
        if mp_world > 1:
            t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
            mp_dist.all_reduce(t, op=mp_dist.ReduceOp.MAX)
            ms, ms_e2e = float(t[0]), float(t[1])
            extra["parallelism"] = "images sharded round-robin over %d ranks, no data-path collective" % mp_world
        tflop = 0.0
        metric, unit = "depth_metrics_img_per_s", "images/s (375x1242 uint8 pairs, 7 metrics each)"
        workload = "my_eval.py depth metrics over 697 synthetic 375x1242 KITTI Eigen-split pairs (BASELINE configs[4])"
        h2d, d2h = int(2 * n_img * h * w), 697 * 8 * 8
        extra["images_per_step"] = n_img
        extra["algorithmic_gb_per_s"] = 2.0 * (n_img / mp_world) * h * w / (ms * 1e-3) / 1e9   # per GPU (one launch)
        extra["frac_of_hbm_peak"] = extra["algorithmic_gb_per_s"] / peaks["hbm"]
    else:
        raise SystemExit("unknown workload " + wl)
    launches = lib.cdb_launch_count() - launches0
    if wl in ("model5", "segcycle") and not args.no_cuda_graph:
        # replays bypass the library's host entry points: kernels of one captured step x timed steps
        launches = int(model._step_graph.launches) * (2 * args.steps + 1)
    if wl == "g_infer" and not args.no_cuda_graph:
        launches = launches + int(getattr(model, "_graph_launches", 0)) * (2 * args.steps)
    clocks = sampler.stop()
    per_step = extra.get("images_per_step", 1)
    if wl == "metrics":
        roof = {"bound": "hbm", "achieved": extra["algorithmic_gb_per_s"], "peak": peaks["hbm"], "unit": "GB/s",
                "frac": extra["frac_of_hbm_peak"], "traffic": int(2 * n_img * h * w + 4.1e6 * n_img / 128),
                "kernel": "metrics_hist_kernel", "us_per_launch": ms * 1e3,
                "algorithmic_bytes_per_unit": "2 B/pixel (gt u8 + pred u8 read once)",
                "peak_source": peaks["source"] + "; traffic from profiles/r01_ncu_full_kernels_r01c.json (128 images: "
                                                 "119.3 MB read + 4.1 MB written), scaled to 697 images"}
    else:
        # tensor-bound steps: the dominant kernel is the same tcgen05 implicit-GEMM family as the headline bench;
        # here the WHOLE step is put against the sustained bf16 peak (kernel timed inside a long step)
        roof = {"bound": "tensor", "achieved": tflop / (ms * 1e-3), "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": tflop / (ms * 1e-3) / peaks["bf16_sustained"], "traffic": None,
                "kernel": "whole step (igemm_flat_kernel / igemm_kernel / wgrad_kernel dominate, "
                          "profiles/r01_step_kernels_*.txt)",
                "peak_source": peaks["source"] + " (sustained)"}
    if mp_world > 1:
        torch.cuda.synchronize()
        mp_dist.barrier()
        if mp_rank != 0:
            mp_dist.destroy_process_group()
            return
    cpu = None if (args.no_cpu_baseline or mp_world > 1) else secondary_cpu_baseline(
        wl, b if wl in ("pix2pix", "model5", "segcycle") else 1)
    line = {
        "metric": metric, "value": per_step * 1e3 / ms, "unit": unit, "n_gpus": mp_world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if mp_world > 1 else "weak", "vs_baseline": None,
        "dtype": "u8/f64" if wl == "metrics" else "bf16", "data": "synthetic",
        "config": dict({"workload": workload, "l2": "inputs + saved activations exceed the 126 MB L2; no explicit flush",
                        "algorithmic_tflop_per_step": tflop,
                        "step_tflops": tflop / (ms * 1e-3) if tflop else None,
                        "step_frac_of_sustained_bf16_peak": tflop / (ms * 1e-3) / peaks["bf16_sustained"] if tflop else None,
                        "device_abort_flag": lib.cdb_device_abort_flag()}, **extra),
        "e2e": {"value": per_step * 1e3 / ms_e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if mp_world > 1:
        mp_dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=256, help="image height (and width unless --width is given)")
    ap.add_argument("--width", type=int, default=0, help="image width for non-square shapes, e.g. --size 192 --width 640")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cudnn-baseline", action="store_true",
                    help="skip config.cudnn_same_box (the stock-PyTorch / cuDNN step timed on this GPU)")
    ap.add_argument("--secondary", action="store_true",
                    help="(default at --gpus 1) also run the other BASELINE configs as sub-runs and attach their lines as "
                         "`secondary`: ~10 s each")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary sub-runs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch per GPU (default); strong: --batch is the global batch, sharded over the ranks")
    ap.add_argument("--no-cuda-graph", action="store_true", help="eager launches instead of replaying the captured step")
    ap.add_argument("--no-batch-passes", action="store_true",
                    help="run every generator / discriminator pass separately (6 + 2 per update) instead of batching "
                         "the passes that share a network")
    ap.add_argument("--workload", default="cyclegan", choices=["cyclegan", "pix2pix", "model5", "g_infer", "metrics", "segcycle"],
                    help="cyclegan (default, the headline metric) or one of the secondary BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload != "cyclegan":
        secondary_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
