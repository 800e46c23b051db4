import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
names = [n for n, _, _ in synth.CONV_LAYERS]
W = synth.vgg_weights(seed=5)
a = vgg.StyleContentModel(names[:1], names[1:], weights=W)
b = vgg.StyleContentModel(names[:1], names[1:], weights=W); b.vgg.set_conv_path("simt")
for (H, Wd) in [(48, 80), (64, 64), (48, 69)]:
    img = torch.as_tensor(synth.image(H, Wd, 0)).cuda()
    oa, ob = a(img), b(img)
    fa = dict(oa["content"]); fa.update(oa["style"]); fb = dict(ob["content"]); fb.update(ob["style"])
    for n in names:
        e = (fa[n] - fb[n]).abs()
        rel = float(e.max() / fb[n].abs().max())
        print(H, Wd, n, tuple(fa[n].shape), "rel %.2e" % rel, "" if rel < 1e-5 else " <<<< BAD at %s" % (np.unravel_index(int(e.argmax()), e.shape),))
    rng = np.random.default_rng(1)
    seeds = {n: torch.as_tensor(rng.standard_normal(tuple(fa[n].shape)).astype(np.float32)).cuda() for n in ("block4_conv2", "block1_conv1", "block2_conv1", "block3_conv1", "block4_conv1", "block5_conv1")}
    ga, gb = a.backward(seeds), b.backward(seeds)
    print(H, Wd, "backward rel %.2e" % float((ga - gb).abs().max() / gb.abs().max()))
