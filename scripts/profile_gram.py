"""Profiling driver: the masked Gram of one style layer at the benchmark shape (default block1_conv1: 1024x1024x64, K = 8)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = "automated-deep-photo-style-transfer_b200"
k = importlib.import_module(pkg + ".kernels"); synth = importlib.import_module(pkg + ".synth")
sem = importlib.import_module(pkg + ".components.semantic_merge")
hw, C, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 8
F = (torch.rand(hw, hw, C, device="cuda") * 50).contiguous()
m = torch.stack([t[0, :, :, 0] for t in sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(1024, 1024, K, 9)))]).cuda()
m = torch.nn.functional.interpolate(m[None], size=(hw, hw), mode="bilinear", align_corners=False)[0].reshape(K, hw * hw).contiguous()
pl = k.gram_patch_lists(m, hw, hw, K, "cuda")
ws = k.gram_workspace(hw * hw, C, K, "cuda")
fa, ma = k.absmax_slot(F), k.absmax_slot(m)
for _ in range(4):
    G = k.gram_masked(F, m, K, ws, patches=pl, f_absmax=fa.data_ptr(), masks_absmax=ma)
torch.cuda.synchronize(); print("ok", float(G.abs().max()))
