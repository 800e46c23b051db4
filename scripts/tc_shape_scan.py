import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lib = importlib.import_module(pkg + "._lib"); L = lib.lib()
names = [n for n, _, _ in synth.CONV_LAYERS]
ext = vgg.StyleContentModel(names[:1], names[1:], weights=synth.vgg_weights(seed=7))
def run(fn, i, x, h, w, c, path):
    lib.check(L.adpst_vgg_set_conv_path(ext.vgg._h, path))
    y = torch.full((h, w, c), float("nan"), dtype=torch.float32, device="cuda")
    lib.check(getattr(L, fn)(ext.vgg._h, i, lib.ptr(x), h, w, lib.ptr(y), None, lib.stream_ptr())); torch.cuda.synchronize()
    lib.check(L.adpst_vgg_set_conv_path(ext.vgg._h, 0)); return y
bad = 0
for (h, w) in [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5), (48, 69), (24, 34), (12, 17), (6, 8), (3, 4), (8, 16), (9, 17), (16, 32)]:
    for i in (1, 2, 4, 8, 12):
        cin, cout = synth.CONV_LAYERS[i][1], synth.CONV_LAYERS[i][2]
        g = torch.Generator(device="cuda").manual_seed(i + h)
        x = (torch.rand(h, w, cin, device="cuda", generator=g) * 200).contiguous()
        d = torch.randn(h, w, cout, device="cuda", generator=g).contiguous()
        e1 = (run("adpst_vgg_conv_forward", i, x, h, w, cout, 0) - run("adpst_vgg_conv_forward", i, x, h, w, cout, 1))
        r1 = run("adpst_vgg_conv_forward", i, x, h, w, cout, 1)
        e2 = (run("adpst_vgg_conv_dgrad", i, d, h, w, cin, 0) - run("adpst_vgg_conv_dgrad", i, d, h, w, cin, 1))
        r2 = run("adpst_vgg_conv_dgrad", i, d, h, w, cin, 1)
        a, b = float(e1.abs().max() / r1.abs().max()), float(e2.abs().max() / r2.abs().max())
        flag = "" if (a < 1e-5 and b < 1e-5) else "  <<<<<< BAD"
        if flag: 
            bad += 1
            idx = (e2.abs() if b >= 1e-5 else e1.abs()).amax(dim=2)
            print("   worst pixels (y,x):", [(int(k // w), int(k % w)) for k in torch.topk(idx.flatten(), 5).indices])
        print("h=%3d w=%3d layer %2d fwd %.2e dgrad %.2e%s" % (h, w, i, a, b, flag))
print("bad:", bad)
