"""Development helper: time one train_step (eager and CUDA-graph) at a given size."""
import argparse, importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth")
st = importlib.import_module(pkg + ".style_transfer")
vgg = importlib.import_module(pkg + ".components.VGG19.model")
lossm = importlib.import_module(pkg + ".components.loss")
sem = importlib.import_module(pkg + ".components.semantic_merge")
ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=512); ap.add_argument("--K", type=int, default=4)
ap.add_argument("--iters", type=int, default=5); a = ap.parse_args()
H = W = a.size
args = argparse.Namespace(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=1e4, matting_epsilon=1e-7,
                          matting_window_radius=1, adam_lr=0.1, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8)
w = synth.vgg_weights()
c = torch.as_tensor(synth.image(H, W, 0)).cuda(); s = torch.as_tensor(synth.image(H, W, 1)).cuda()
cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, a.K, 9)))
sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, a.K, 10)))
ext = vgg.StyleContentModel(st.CONTENT_LAYERS, st.STYLE_LAYERS, weights=w)
loss = lossm.Loss(ext(c)["content"], ext(s)["style"], args, cm, sm)
loss.initialize_matting_laplacian(c[0].double())
for graph in (False, True):
    opt = st.Adam(); x = c.clone()
    step = st.make_train_step(ext, loss, opt, use_cuda_graph=graph)
    for _ in range(2): step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): d = step(x)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"size": a.size, "K": a.K, "graph": graph, "ms_per_step": e0.elapsed_time(e1) / a.iters, "total": float(d["Total loss"])}))
# per-phase timing (eager)
def tm(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
x = c.clone()
print("forward ms", tm(lambda: ext(x, reuse=True)))
outs = ext(x, reuse=True)
print("loss ms", tm(lambda: loss(x, outs)))
print("backward ms", tm(lambda: loss.gradient(ext)))
