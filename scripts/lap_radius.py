import importlib, sys, os
sys.path.insert(0, "/root/repo")
import torch
pkg = "automated-deep-photo-style-transfer_b200"
v2 = importlib.import_module(pkg + ".components.matting_v2")
synth = importlib.import_module(pkg + ".synth")
for size in (512, 1024):
    img = torch.as_tensor(synth.image(size, size, 0)[0]).cuda()
    x = torch.rand(size * size, 3, device="cuda"); y = torch.empty_like(x)
    for r in (1, 2, 3):
        op = v2.MattingLaplacian(img, epsilon=1e-5, window_radius=r, storage_dtype=torch.float32, compute_dtype=torch.float64)
        fn = lambda: op._op.apply3(x, want_y=True, want_quad=True, y_scale=2.0, out=y)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        print("size %d r=%d: %.3f ms" % (size, r, a.elapsed_time(b) / 10), flush=True)
