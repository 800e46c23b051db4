import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lossm = importlib.import_module(pkg + ".components.loss"); sem = importlib.import_module(pkg + ".components.semantic_merge")
st = importlib.import_module(pkg + ".style_transfer")
H, Wd, K = 48, 80, 4
args = argparse.Namespace(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=0.0, matting_epsilon=1e-7,
                          matting_window_radius=1)
W = synth.vgg_weights(seed=5)
content, style = synth.image(H, Wd, 0), synth.image(H, Wd, 1)
cell = 12
cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, Wd, K, 9, cell=cell)))
sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, Wd, K, 10, cell=cell)))
res = {}
for path in ("tensor", "simt"):
    ext = vgg.StyleContentModel(st.CONTENT_LAYERS, st.STYLE_LAYERS, weights=W); ext.vgg.set_conv_path(path)
    c, s = torch.as_tensor(content).cuda(), torch.as_tensor(style).cuda()
    loss = lossm.Loss(ext(c)["content"], ext(s)["style"], args, cm, sm)
    pert = np.sign(synth.image(H, Wd, 3) - 0.5).astype(np.float32) * 0.1
    x = torch.clamp(c + torch.as_tensor(pert).cuda(), 0, 1).contiguous()
    d = loss(x, ext(x, reuse=True))
    seeds = {k: v.clone() for k, v in loss._seeds.items()}
    g = loss.gradient(ext).clone()
    res[path] = ({k: float(v) for k, v in d.items()}, seeds, g, {k: v["A"].clone() for k, v in loss._layer_cache.items()})
print(res["tensor"][0]); print(res["simt"][0])
for k in res["simt"][1]:
    a, b = res["tensor"][1][k], res["simt"][1][k]
    print("seed", k, "rel %.2e" % float((a - b).abs().max() / b.abs().max()))
for k in res["simt"][3]:
    a, b = res["tensor"][3][k], res["simt"][3][k]
    print("styleGram", k, "rel %.2e" % float((a - b).abs().max() / b.abs().max()))
a, b = res["tensor"][2], res["simt"][2]
print("grad rel %.2e" % float((a - b).abs().max() / b.abs().max()))
d = (a - b).abs()[0].amax(dim=2)
thr = 1e-5 * float(b.abs().max())
ys, xs = torch.nonzero(d > thr, as_tuple=True)
print("pixels over 1e-5*max: %d of %d; bbox y[%d,%d] x[%d,%d]" % (len(ys), d.numel(), int(ys.min()), int(ys.max()), int(xs.min()), int(xs.max())))
# ReLU sign flips / pool argmax flips between the two paths
names = [n for n, _, _ in synth.CONV_LAYERS]
ea = vgg.StyleContentModel(names[:1], names[1:], weights=W); eb = vgg.StyleContentModel(names[:1], names[1:], weights=W); eb.vgg.set_conv_path("simt")
oa, ob = ea(x), eb(x)
fa = dict(oa["content"]); fa.update(oa["style"]); fb = dict(ob["content"]); fb.update(ob["style"])
import torch.nn.functional as F
for n in names:
    A_, B_ = fa[n], fb[n]
    flips = int(((A_ > 0) != (B_ > 0)).sum())
    pa = F.max_pool2d(A_.permute(0, 3, 1, 2), 2, 2, return_indices=True)[1]; pb = F.max_pool2d(B_.permute(0, 3, 1, 2), 2, 2, return_indices=True)[1]
    print(n, "relu flips", flips, "pool argmax flips", int((pa != pb).sum()), "of", pa.numel())
