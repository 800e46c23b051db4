"""Profiling driver: one tensor-core conv layer (block3_conv2: 256->256 at 256x256) forward, a few launches."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lib = importlib.import_module(pkg + "._lib"); L = lib.lib()
names = [n for n, _, _ in synth.CONV_LAYERS]
ext = vgg.StyleContentModel(names[:1], names[1:], weights=synth.vgg_weights())
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 5
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cin, cout = synth.CONV_LAYERS[layer][1], synth.CONV_LAYERS[layer][2]
x = torch.rand(hw, hw, cin, device="cuda") * 100
y = torch.empty(hw, hw, cout, device="cuda")
for _ in range(3):
    lib.check(L.adpst_vgg_conv_forward(ext.vgg._h, layer, lib.ptr(x), hw, hw, lib.ptr(y), None, lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    lib.check(L.adpst_vgg_conv_forward(ext.vgg._h, layer, lib.ptr(x), hw, hw, lib.ptr(y), None, lib.stream_ptr()))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("layer %d %dx%d: %.3f ms, %.1f TFLOP/s" % (layer, hw, hw, ms, 2.0 * hw * hw * 9 * cin * cout / ms / 1e9))
