import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
v2 = importlib.import_module(pkg + ".components.matting_v2"); synth = importlib.import_module(pkg + ".synth")
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cd = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
img = torch.as_tensor(synth.image(size, size, 0)[0]).cuda(); x = torch.rand(size * size, 3, device="cuda"); y = torch.empty_like(x)
op = v2.MattingLaplacian(img, epsilon=1e-7, storage_dtype=torch.float32, compute_dtype=cd)
for _ in range(4): op._op.apply3(x, want_y=True, want_quad=True, y_scale=2.0, out=y)
torch.cuda.synchronize(); print("ok")
