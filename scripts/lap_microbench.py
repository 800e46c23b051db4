"""Quick L.x microbenchmark (CUDA events, L2 flushed between timed launches).  Development helper."""
import importlib, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
v2 = importlib.import_module(pkg + ".components.matting_v2")
v3 = importlib.import_module(pkg + ".components.matting_v3")
synth = importlib.import_module(pkg + ".synth")

def time_op(fn, iters=20, flush=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2], ts[0]

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for size in (256, 512, 1024, 2048, 4096):
    img = torch.as_tensor(synth.image(size, size, 0)[0]).cuda()
    x = torch.rand(size*size, 3, device="cuda")
    y = torch.empty_like(x)
    for name, cls, cd in (("v2/f64", v2, torch.float64), ("v2/f32", v2, torch.float32), ("v3/f64", v3, torch.float64), ("v3/f32", v3, torch.float32)):
        op = cls.MattingLaplacian(img, epsilon=1e-7, storage_dtype=torch.float32, compute_dtype=cd)
        med, best = time_op(lambda: op._op.apply3(x, want_y=True, want_quad=True, y_scale=2.0, out=y), flush=flush)
        gbs = 36.0 * size * size / (med * 1e-3) / 1e9
        print(json.dumps({"size": size, "variant": name, "ms_median": round(med, 4), "ms_best": round(best, 4), "GBps_algorithmic": round(gbs, 1)}))
