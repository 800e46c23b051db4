import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
k = importlib.import_module("automated-deep-photo-style-transfer_b200.kernels")
torch.manual_seed(0)
for (h, w, C, K) in [(2, 16, 128, 1), (8, 16, 128, 1), (8, 32, 64, 1), (8, 32, 256, 1), (16, 32, 128, 2)]:
    F = (torch.rand(h, w, C, device="cuda") * 10).contiguous()
    m = None
    if K > 1:
        lab = torch.randint(0, K, (h * w,), device="cuda")
        m = torch.stack([(lab == i).float() for i in range(K)]).contiguous()
    pl = k.gram_patch_lists(m, h, w, K, "cuda")
    Gt = k.gram_masked(F, m, K, path="tensor", patches=pl)
    Gs = k.gram_masked(F, m, K, path="simt")
    X = F.reshape(-1, C).double()
    ref = torch.stack([(X * (m[i].double()[:, None] if m is not None else 1)).T @ (X * (m[i].double()[:, None] if m is not None else 1)) for i in range(K)])
    print((h, w, C, K), "patches", pl[0].tolist()[:8], pl[1].tolist(), "tc err %.2e simt err %.2e" % (float((Gt.double() - ref).abs().max() / ref.abs().max()), float((Gs.double() - ref).abs().max() / ref.abs().max())))
    if float((Gt.double() - ref).abs().max() / ref.abs().max()) > 1e-4:
        print("  Gt[0,:4,:4]\n", Gt[0, :4, :4].cpu().numpy(), "\n  ref\n", ref[0, :4, :4].cpu().numpy())
        r = (Gt[0].double() / ref[0])
        print("  ratio stats: min %.3f max %.3f; nonzero frac %.3f" % (float(r.min()), float(r.max()), float((Gt[0] != 0).float().mean())))
