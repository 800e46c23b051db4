import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
k = importlib.import_module("automated-deep-photo-style-transfer_b200.kernels")
n = 12288
rng = np.random.default_rng(n)
x0 = rng.random(n).astype(np.float32)
x = torch.as_tensor(x0).cuda(); st = k.AdamState(x)
f = np.float32; b2 = f(0.999); om2 = f(1) - b2; b1 = f(0.9); om1 = f(1) - b1
v32 = np.zeros(n, f); m32 = np.zeros(n, f)
for t in range(1, 4):
    sc = 10.0 ** rng.integers(-3, 3)
    g = (rng.standard_normal(n) * sc).astype(f)
    k.adam_clip_step(x, torch.as_tensor(g).cuda(), st, 0.1, 0.9, 0.999, 1e-8)
    v32 = b2 * v32 + (om2 * g) * g
    m32 = b1 * m32 + om1 * g
    gv = st.v.cpu().numpy(); gm = st.m.cpu().numpy()
    bad = np.nonzero(np.abs(gv - v32) > 1e-5 * np.abs(v32) + 1e-12)[0]
    print("t", t, "scale", sc, "bad v", len(bad), "bad m", int((np.abs(gm - m32) > 1e-5 * np.abs(m32) + 1e-12).sum()))
    for i in bad[:8]:
        print("  i", i, "g", g[i], "gpu v", gv[i], "cpu v", v32[i], "gpu m", gm[i], "cpu m", m32[i])
