"""Profiling driver: one tensor-core conv layer, forward (dir=0) or data gradient (dir=1), a few launches."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lib = importlib.import_module(pkg + "._lib"); L = lib.lib()
names = [n for n, _, _ in synth.CONV_LAYERS]
ext = vgg.StyleContentModel(names[:1], names[1:], weights=synth.vgg_weights())
layer, hw, grad = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0
cin, cout = synth.CONV_LAYERS[layer][1], synth.CONV_LAYERS[layer][2]
K, N = (cout, cin) if grad else (cin, cout)
x = torch.rand(hw, hw, K, device="cuda") * 100
y = torch.empty(hw, hw, N, device="cuda")
slot = torch.zeros(1, dtype=torch.int32, device="cuda")
lib.check(L.adpst_absmax(lib.ptr(x), x.numel(), lib.ptr(slot), lib.stream_ptr()))
fn = L.adpst_vgg_conv_dgrad if grad else L.adpst_vgg_conv_forward
for _ in range(5):
    lib.check(fn(ext.vgg._h, layer, lib.ptr(x), hw, hw, lib.ptr(y), lib.ptr(slot), lib.stream_ptr()))
torch.cuda.synchronize(); print("ok")
