import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
np.set_printoptions(linewidth=200)
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lib = importlib.import_module(pkg + "._lib"); L = lib.lib()
names = [n for n, _, _ in synth.CONV_LAYERS]
ext = vgg.StyleContentModel(names[:1], names[1:], weights=synth.vgg_weights())
layer, hw, blk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cin, cout = synth.CONV_LAYERS[layer][1], synth.CONV_LAYERS[layer][2]
x = torch.rand(hw, hw, cin, device="cuda") * 100; y = torch.empty(hw, hw, cout, device="cuda")
buf = torch.zeros(5 * 4096, dtype=torch.int64, device="cuda")
for _ in range(2): lib.check(L.adpst_vgg_conv_forward(ext.vgg._h, layer, lib.ptr(x), hw, hw, lib.ptr(y), None, lib.stream_ptr()))
lib.check(L.adpst_debug_conv_trace(lib.ptr(buf), blk))
lib.check(L.adpst_vgg_conv_forward(ext.vgg._h, layer, lib.ptr(x), hw, hw, lib.ptr(y), None, lib.stream_ptr())); torch.cuda.synchronize()
lib.check(L.adpst_debug_conv_trace(None, -1))
b = buf.cpu().numpy()
t0, t1 = b[4 * 4096], b[4 * 4096 + 1]
iters = 9 * cin // 32
print("CTA total %d clk; iters %d; per iter %.0f" % (t1 - t0, iters, (t1 - t0) / iters))
mm = b[2 * 4096:2 * 4096 + 4 * iters].reshape(iters, 4) - t0
sl = slice(24, 40)
print("mma: arrive      ", mm[sl, 0]); print("     ready seen  ", mm[sl, 1]); print("     issued      ", mm[sl, 2])
print("wait %s\nissue %s\nloop gap %s" % (mm[sl, 1] - mm[sl, 0], mm[sl, 2] - mm[sl, 1], mm[1:, 0][sl] - mm[:-1, 2][sl]))
