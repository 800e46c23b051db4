"""L.x timing of the float64-arithmetic path (what Loss uses) at a few sizes; L2 flushed between launches."""
import importlib, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
v2 = importlib.import_module(pkg + ".components.matting_v2")
v3 = importlib.import_module(pkg + ".components.matting_v3")
synth = importlib.import_module(pkg + ".synth")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]
for size in sizes:
    img = torch.as_tensor(synth.image(size, size, 0)[0]).cuda()
    x = torch.rand(size * size, 3, device="cuda")
    y = torch.empty_like(x)
    for name, cls in (("v2", v2), ("v3", v3)):
        op = cls.MattingLaplacian(img, epsilon=1e-7, storage_dtype=torch.float32, compute_dtype=torch.float64)
        fn = lambda: op._op.apply3(x, want_y=True, want_quad=True, y_scale=2.0, out=y)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort(); med = ts[len(ts) // 2]
        print("%s %d: %.4f ms  %.0f GB/s algorithmic" % (name, size, med, 36.0 * size * size / med / 1e6), flush=True)
