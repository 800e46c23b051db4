"""Profiling driver: a few eager train_steps at 1024x1024, K=8 (same workload as bench.py) -- run under ncu."""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); st = importlib.import_module(pkg + ".style_transfer")
vgg = importlib.import_module(pkg + ".components.VGG19.model"); lossm = importlib.import_module(pkg + ".components.loss")
sem = importlib.import_module(pkg + ".components.semantic_merge")
ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=1024); ap.add_argument("--K", type=int, default=8)
ap.add_argument("--steps", type=int, default=2); a = ap.parse_args()
S, K = a.size, a.K
args = argparse.Namespace(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=1e4, matting_epsilon=1e-7,
                          matting_window_radius=1, adam_lr=0.1, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, tv_weight=1.0)
c = torch.as_tensor(synth.image(S, S, 0)).cuda(); s = torch.as_tensor(synth.image(S, S, 1)).cuda()
cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(S, S, K, 9)))
sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(S, S, K, 10)))
ext = vgg.StyleContentModel(st.CONTENT_LAYERS, st.STYLE_LAYERS, weights=synth.vgg_weights())
loss = lossm.Loss(ext(c)["content"], ext(s)["style"], args, cm, sm)
loss.initialize_matting_laplacian(c[0].double())
step = st.make_train_step(ext, loss, st.Adam())
x = c.clone()
step(x); torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    d = step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("total", float(d["Total loss"]))
