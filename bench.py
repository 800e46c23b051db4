#!/usr/bin/env python
"""bench.py -- Adam iterations/second of the style-transfer hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one full train_step (style_transfer.py:331-344 of the reference): VGG19 forward to block5_conv1,
content + masked-Gram style + photorealism (x^T L x) losses, gradient to the image, Adam + clip.
Workload at every N: BASELINE.json configs[1] -- one 1024x1024 content/style pair per GPU, 8 semantic classes,
matting_v2 Laplacian (eps 1e-7, r 1), synthetic U[0,1) images, seeded He-normal VGG19 weights.
N > 1: independent pairs, one per rank, no data-path collective -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value` is device-timed with the inputs resident in HBM (CUDA-graph replay);
`e2e` is the same metric through the public train_step API with the image starting and ending in pinned HOST memory
every step.  `roofline` describes the dominant kernel family (3x3 conv, tensor-bound) and `roofline_lx` the
Laplacian mat-vec (HBM-bound).  `cpu_baseline` (rank 0, N = 1) times the oracle port on the host cores.

--impl reference: the reference program itself needs TensorFlow/Keras (absent, no network), so the reference arm
times the CPU port of the same path (oracle/: torch-CPU float32 VGG/Gram + float64 matting_v2), kind "port".
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "automated-deep-photo-style-transfer_b200"

METRIC = "adam_iters_per_sec_1024x1024"
UNIT = "iter/s"


def hyper(size_classes):
    return argparse.Namespace(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=1e4,
                              matting_epsilon=1e-7, matting_window_radius=1, adam_lr=0.1, adam_beta1=0.9,
                              adam_beta2=0.999, adam_epsilon=1e-8)


def conv_flops(size):
    """Algorithmic FLOPs of the 13 forward convolutions (2*h*w*9*Cin*Cout) and the 12 data gradients."""
    synth = importlib.import_module(PKG + ".synth")
    fwd, bwd, h = [], [], size
    i = 0
    for item in synth.VGG_TOPOLOGY:
        if item == "P":
            h //= 2
            continue
        _, cin, cout = item
        f = 2.0 * h * h * 9 * cin * cout
        fwd.append((i, h, cin, cout, f))
        if i > 0:
            bwd.append((i, h, cin, cout, f))
        i += 1
    return fwd, bwd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ----------------------------------------------------------------------------------------------------------------
# CPU port (oracle) timing: cpu_baseline leg and --impl reference
# ----------------------------------------------------------------------------------------------------------------
def cpu_port_state(size, K):
    import torch
    from oracle import masks as omasks
    from oracle import model as omodel
    synth = importlib.import_module(PKG + ".synth")
    a = hyper(None)
    cfg = {"weights": {"content": a.content_weight, "style": a.style_weight, "nima": 0.0, "photo": a.regularization_weight},
           "matting_epsilon": a.matting_epsilon, "matting_window_radius": a.matting_window_radius,
           "adam": {"lr": a.adam_lr, "beta1": a.adam_beta1, "beta2": a.adam_beta2, "epsilon": a.adam_epsilon}}
    cm = [torch.as_tensor(m) for m in omasks.mask_for_tf(omasks.extract_segmentation_masks(synth.label_image(size, size, K, 9)))]
    sm = [torch.as_tensor(m) for m in omasks.mask_for_tf(omasks.extract_segmentation_masks(synth.label_image(size, size, K, 10)))]
    # float32 VGG / Gram (what TF-CPU would run), float64 matting Laplacian (what the reference runs)
    return omodel.TrainState(torch.as_tensor(synth.image(size, size, 0)), torch.as_tensor(synth.image(size, size, 1)),
                             synth.vgg_weights(), cfg, cm, sm, dtype=torch.float32)


def run_reference(args):
    """--impl reference: the CPU port of the same step on all host threads torch will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    st = cpu_port_state(args.size, args.classes)
    budget = float(os.environ.get("ADPST_REF_BUDGET_S", "240"))
    t0 = time.perf_counter()
    st.train_step()
    first = time.perf_counter() - t0
    warm = max(0, min(args.warmup - 1, int(budget * 0.2 / max(first, 1e-3))))
    for _ in range(warm):
        st.train_step()
    n = max(1, min(args.steps, int(budget * 0.8 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        st.train_step()
    dt = time.perf_counter() - t0
    v = n / dt
    cores = torch.get_num_threads()
    sample = "%d of the requested %d iterations at %dx%d, K=%d (time-bounded to %.0f s); %d warm-up" % (
        n, args.steps, args.size, args.size, args.classes, budget, warm + 1)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "steps_executed": n, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (VGG/Gram) + f64 (Laplacian)", "data": "synthetic",
        "config": {"workload": "1x %dx%d pair, %d classes, matting_v2 eps=1e-7 r=1, CPU port of the reference step"
                               % (args.size, args.size, args.classes)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "TensorFlow/Keras are not installable here (no network): this is the oracle port, not the TF program",
    }))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def time_launches(fn, flush, reps=5):
    import torch
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def time_rotating(fns, rounds=5):
    """Per-call device time of a set of equivalent calls whose combined working set exceeds the 126 MB L2: the calls are
    queued back to back (no host launch gap inside the timed region) and each one finds its inputs evicted."""
    import torch
    for f in fns:
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rounds):
        for f in fns:
            f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (rounds * len(fns))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    synth = importlib.import_module(PKG + ".synth")
    st = importlib.import_module(PKG + ".style_transfer")
    vggm = importlib.import_module(PKG + ".components.VGG19.model")
    lossm = importlib.import_module(PKG + ".components.loss")
    sem = importlib.import_module(PKG + ".components.semantic_merge")
    v2 = importlib.import_module(PKG + ".components.matting_v2")
    lib = importlib.import_module(PKG + "._lib")

    S, K = args.size, args.classes
    hp = hyper(None)
    content_h = torch.as_tensor(synth.image(S, S, 2 * rank)).pin_memory()
    style = torch.as_tensor(synth.image(S, S, 2 * rank + 1)).cuda()
    content = content_h.cuda()
    cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(S, S, K, 9 + rank)))
    sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(S, S, K, 10 + rank)))
    ext = vggm.StyleContentModel(st.CONTENT_LAYERS, st.STYLE_LAYERS, shape=(None, None, 3), weights=synth.vgg_weights())
    loss = lossm.Loss(ext(content)["content"], ext(style)["style"], hp, cm, sm)
    loss.initialize_matting_laplacian(content[0].to(torch.float64))
    opt = st.Adam(hp.adam_lr, hp.adam_beta1, hp.adam_beta2, hp.adam_epsilon)
    x = content.clone()

    # launches per step (eager), then the replayable graph
    eager = st.make_train_step(ext, loss, opt, use_cuda_graph=False)
    eager(x); torch.cuda.synchronize()
    n0 = lib.launch_count(); eager(x); torch.cuda.synchronize()
    launches_per_step = lib.launch_count() - n0
    x.copy_(content); opt._slots.m.zero_(); opt._slots.v.zero_(); opt._slots.state.zero_()
    step = st.make_train_step(ext, loss, opt, use_cuda_graph=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(x)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        d = step(x)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    clocks = sampler.stop() if rank == 0 else None
    final_total = float(d["Total loss"])

    # e2e: image starts and ends in pinned host memory every step; losses read back every step
    host_img = content_h.clone().pin_memory()
    host_loss = torch.empty(5, dtype=torch.float32).pin_memory()
    x.copy_(content); opt._slots.m.zero_(); opt._slots.v.zero_(); opt._slots.state.zero_()
    for _ in range(3):
        x.copy_(host_img, non_blocking=True); step(x); host_img.copy_(x, non_blocking=True)
        host_loss.copy_(loss._out, non_blocking=True); torch.cuda.current_stream().synchronize()
    barrier()
    t0 = time.perf_counter(); e0.record()
    for _ in range(args.steps):
        x.copy_(host_img, non_blocking=True)
        step(x)
        host_img.copy_(x, non_blocking=True)
        host_loss.copy_(loss._out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record(); barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2)
    wall_e2e = time.perf_counter() - t0

    out = None
    if rank == 0:
        hbm, tf_burst, tf_sus, which = peaks()
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
        # --- dominant kernel family: 3x3 convolutions, one launch per layer, timed alone with L2 flushed
        fwd, bwd = conv_flops(S)
        A = ext._loop
        tot_f, tot_t, n_launch = 0.0, 0.0, 0
        scratch = torch.empty(A.acts[0].numel(), dtype=torch.float32, device="cuda")
        stream = lib.stream_ptr()
        for i, h, cin, cout, f in fwd:
            src = x if i == 0 else (A.pools[[1, 3, 7, 11].index(i - 1)] if (i - 1) in (1, 3, 7, 11) else A.acts[i - 1])
            # the input's scale slot is the one the step's own forward pass left behind (max|conv i-1 output|)
            slot = ext.vgg.act_absmax_ptr(i - 1) if i > 0 else None
            t = time_launches(lambda: lib.check(lib.lib().adpst_vgg_conv_forward(ext.vgg._h, i, lib.ptr(src), h, h,
                                                                                 lib.ptr(scratch), slot, stream)), flush, 3)
            tot_f += f; tot_t += t; n_launch += 1
        for i, h, cin, cout, f in bwd:
            slot = ext.vgg.act_absmax_ptr(i)
            t = time_launches(lambda: lib.check(lib.lib().adpst_vgg_conv_dgrad(ext.vgg._h, i, lib.ptr(A.acts[i]), h, h,
                                                                               lib.ptr(scratch), slot, stream)), flush, 3)
            tot_f += f; tot_t += t; n_launch += 1
        conv_tflops = tot_f / (tot_t * 1e-3) / 1e12
        # The reference computes these convolutions in float32 (1e-5 parity): the tensor-core kernel forms every product from
        # 3 FP16 MMAs (hi*hi + hi*lo + lo*hi, power-of-two scaled, fp32 accumulation), which run at the bf16 rate, so the
        # ceiling for this arithmetic is peak_bf16 / 3; `frac` is still quoted against the measured bf16 peak, as the
        # contract asks.
        roofline = {"kernel": "conv3x3_tc_kernel: persistent tcgen05 3xFP16 implicit GEMM, TMA-fed, A operand in TMEM (12 forward "
                              "+ 12 data-gradient launches per step; block1_conv1 forward is a CUDA-core kernel)", "bound": "tensor",
                    "achieved": conv_tflops, "peak": tf_sus, "unit": "TFLOP/s", "frac": conv_tflops / tf_sus,
                    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (block3_conv2 forward: 67 MB in, 67 MB out,
                    # 2.4 MB of weights) from the ncu --set full capture summarised in profiles/r1_prof_conv_summary.csv
                    "traffic": 91.9e6, "traffic_of": "block3_conv2 forward launch, ncu capture in profiles/ (algorithmic: 136.6 MB; "
                                                     "part of the output is still in L2 when the kernel ends)",
                    "peak_source": which + " bf16 dense, sustained (kernel timed inside a long step)",
                    "flops_per_launch_avg": tot_f / n_launch, "ms_per_launch_avg": tot_t / n_launch,
                    "share_of_step": tot_t / (total_ms / args.steps),
                    "fp32_accurate_ceiling": {"what": "3xFP16: peak_bf16 / 3 (MMAs per product)",
                                              "peak": tf_sus / 3.0, "frac": conv_tflops / (tf_sus / 3.0)},
                    "note": "algorithmic FLOPs (2*h*w*9*Cin*Cout per layer); issued tensor FLOPs are 3x that"}
        # --- Laplacian mat-vec (fused x^T L x and 2Lx), 36 B/px algorithmic
        # L2 is kept cold by rotating over enough independent (operator, x, y) sets to exceed it (each set is 36 B/px)
        nsets = max(2, int(160e6 // (36 * S * S)) + 1)
        xs = x.reshape(-1, 3)

        def lap_calls(compute_dtype):
            calls = []
            for j in range(nsets):
                op = v2.MattingLaplacian(torch.roll(content[0], j, 0).contiguous(), epsilon=1e-7, storage_dtype=torch.float32,
                                         compute_dtype=compute_dtype)
                xj, yj, qj = torch.roll(xs, j, 0).contiguous(), torch.empty_like(xs), torch.zeros(1, dtype=torch.float64, device="cuda")
                calls.append(lambda op=op, xj=xj, yj=yj, qj=qj: op._op.apply3(xj, want_y=True, want_quad=True, y_scale=2e4,
                                                                               out=yj, quad_out=qj))
            return calls
        t_lx = time_rotating(lap_calls(torch.float64))
        lx_gbs = 36.0 * S * S / (t_lx * 1e-3) / 1e9
        t_lx32 = time_rotating(lap_calls(torch.float32))
        # The quoted roofline is HBM (36 B/px), as the metric asks; what actually bounds the kernel is the float64 pipe:
        # ~250 float64 lane-operations per pixel (halo included) against 64 lanes/clk/SM.
        f64_floor_ms = 250.0 * S * S / (64.0 * 148 * 1.965e9) * 1e3
        roofline_lx = {"kernel": "lap_march3_kernel (float32 I/O, float64 arithmetic: the path Loss uses)", "bound": "hbm",
                       "achieved": lx_gbs, "peak": hbm, "unit": "GB/s", "frac": lx_gbs / hbm,
                       # dram bytes of one 2048x2048 launch (ncu capture in profiles/r1_prof_lap_summary.csv; 151 MB algorithmic)
                       "traffic": 123.8e6, "traffic_of": "2048x2048 launch, ncu capture in profiles/",
                       "peak_source": which + " copy bandwidth", "bytes_per_launch": 36 * S * S, "ms_per_launch": t_lx,
                       "l2": "%d independent operator/x/y sets (%.0f MB) rotated, calls queued back to back" % (nsets, nsets * 36e-6 * S * S),
                       "float64_pipe_floor": {"what": "250 float64 lane-ops/px at 64 lanes/clk/SM x 148 SMs x 1.965 GHz",
                                              "ms_per_launch": f64_floor_ms, "frac_of_floor": f64_floor_ms / t_lx},
                       "float32_arithmetic_variant": {"achieved": 36.0 * S * S / (t_lx32 * 1e-3) / 1e9,
                                                      "frac": 36.0 * S * S / (t_lx32 * 1e-3) / 1e9 / hbm,
                                                      "ms_per_launch": t_lx32}}
        del flush, scratch
        value = world * args.steps / (total_ms * 1e-3)
        e2e = world * args.steps / (e2e_ms * 1e-3)
        nbytes = x.numel() * 4
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (VGG/Gram: float32-accurate 3xFP16 tensor-core products, fp32 accumulation; Adam f32) + f64 (Laplacian arithmetic)",
            "data": "synthetic",
            "config": {"workload": "configs[1]: one %dx%d content/style pair per GPU, %d semantic classes, content+masked-Gram "
                                   "style+photorealism loss, gradient, Adam+clip; matting_v2 eps=1e-7 r=1; random-init VGG19"
                                   % (S, S, K),
                       "pairs": world, "parallelism": "independent pairs, one per GPU, no collective",
                       "l2": "per-step working set (1.2 GB of activations) exceeds the 126 MB L2; per-kernel timings flush L2",
                       "cuda_graph": True},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 20,
                    "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "roofline_lx": roofline_lx,
            "final_total_loss": final_total,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(S, K)
        print(json.dumps(out))


def cpu_baseline(size, K):
    """Oracle port on the host cores: bounded sample (about 10-30 s of CPU work)."""
    import torch
    st = cpu_port_state(size, K)
    t0 = time.perf_counter(); st.train_step(); first = time.perf_counter() - t0
    n = max(1, min(3, int(20.0 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        st.train_step()
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "host_cpus": os.cpu_count(),
            "sample": "%d iterations at %dx%d, K=%d after 1 warm-up (torch-CPU float32 VGG/Gram + float64 matting_v2 port "
                      "of the reference step; TensorFlow itself is not installable here)" % (n, size, size, K)}


def run_tiled(args):
    """BASELINE configs[3]: ONE 3840x2160 image, column strips over N GPUs (tiled.py): overlapped halos instead of
    per-layer exchange, NCCL all-reduce of the Gram partials, NCCL all-gather of the updated strips.  Strong scaling."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    synth = importlib.import_module(PKG + ".synth"); tiled = importlib.import_module(PKG + ".tiled")
    sem = importlib.import_module(PKG + ".components.semantic_merge")
    H, W, K = args.tiled_h, args.tiled_w, args.classes
    hp = hyper(None)
    content, style = synth.image(H, W, 0), synth.image(H, W, 1)
    cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, K, 9, cell=64)))
    sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, K, 10, cell=64)))
    if world > 1:
        job = tiled.TiledStyleTransfer(content, style, hp, cm, sm, synth.vgg_weights(), rank, world)
    else:
        job = tiled.TiledStyleTransfer(content, style, hp, cm, sm, synth.vgg_weights(), 0, 1, reduce_sum=lambda t: None,
                                       gather=lambda s: [s])
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        d = job.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    for _ in range(args.steps):
        d = job.step()
    e1.record(); barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = job.tile
        print(json.dumps({"metric": "adam_iters_per_sec_%dx%d_spatially_tiled" % (W, H), "value": args.steps / (float(ms) * 1e-3),
                          "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": float(ms) / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32 (float32-accurate 3xFP16 tensor-core products) + f64 (Laplacian arithmetic)", "data": "synthetic",
                          "config": {"workload": "configs[3]: one %dx%d image, %d classes, column strips of %d px + %d px halo per "
                                                 "interior side (local width %d), Gram all-reduce + strip all-gather per step"
                                                 % (W, H, K, W // world, tiled.HALO, t.local_w)},
                          "final_total_loss": float(d["Total loss"])}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--tiled", action="store_true", help="configs[3]: one large image tiled spatially over the GPUs")
    ap.add_argument("--tiled-h", dest="tiled_h", type=int, default=2160)
    ap.add_argument("--tiled-w", dest="tiled_w", type=int, default=3840)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.tiled:
        run_tiled(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
