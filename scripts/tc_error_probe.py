"""Development helper: error of the tcgen05 3xFP16 conv vs the fp32 CUDA-core kernel and vs a float64 torch conv."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lib = importlib.import_module(pkg + "._lib"); L = lib.lib()
W = synth.vgg_weights(seed=7)
names = [n for n, _, _ in synth.CONV_LAYERS]
ext = vgg.StyleContentModel(names[:1], names[1:], weights=W)
def run(fn, i, x, h, w, c, path):
    lib.check(L.adpst_vgg_set_conv_path(ext.vgg._h, path))
    y = torch.empty(h, w, c, dtype=torch.float32, device="cuda")
    lib.check(getattr(L, fn)(ext.vgg._h, i, lib.ptr(x), h, w, lib.ptr(y), None, lib.stream_ptr())); torch.cuda.synchronize()
    lib.check(L.adpst_vgg_set_conv_path(ext.vgg._h, 0)); return y
for i in (1, 2, 4, 8, 9, 12):
    cin, cout = synth.CONV_LAYERS[i][1], synth.CONV_LAYERS[i][2]
    h = w = 32
    g = torch.Generator(device="cuda").manual_seed(i)
    x = (torch.rand(h, w, cin, device="cuda", generator=g) * 200.0).contiguous()
    k, b = W[names[i]]
    ref = F.relu(F.conv2d(x.double().permute(2, 0, 1)[None], torch.as_tensor(k).double().cuda().permute(3, 2, 0, 1),
                          torch.as_tensor(b).double().cuda(), padding=1))[0].permute(1, 2, 0)
    ytc, ysm = run("adpst_vgg_conv_forward", i, x, h, w, cout, 0), run("adpst_vgg_conv_forward", i, x, h, w, cout, 1)
    sc = float(ref.abs().max())
    pos = ref > 0.1 * sc
    print("fwd layer %2d K=%4d  tc-vs-f64 max %.2e  simt-vs-f64 max %.2e  tc signed-mean-rel (big outputs) %.2e  simt %.2e" % (
        i, 9 * cin, float((ytc.double() - ref).abs().max()) / sc, float((ysm.double() - ref).abs().max()) / sc,
        float(((ytc.double() - ref) / ref)[pos].mean()), float(((ysm.double() - ref) / ref)[pos].mean())))
