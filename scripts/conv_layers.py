"""Per-layer timing of the tensor-core conv kernel at the shapes of a S x S image (forward and data gradient),
L2 flushed between launches, input scale slots precomputed (so only the conv kernel is timed)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = "automated-deep-photo-style-transfer_b200"
synth = importlib.import_module(pkg + ".synth"); vgg = importlib.import_module(pkg + ".components.VGG19.model")
lib = importlib.import_module(pkg + "._lib"); L = lib.lib()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
only = [int(a) for a in sys.argv[2:]]
names = [n for n, _, _ in synth.CONV_LAYERS]
ext = vgg.StyleContentModel(names[:1], names[1:], weights=synth.vgg_weights())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
slot = torch.zeros(1, dtype=torch.int32, device="cuda")
pools = {2: 1, 4: 2, 8: 3, 12: 4}
tot_f = tot_t = 0.0
for i in range(1, 13):
    if only and i not in only:
        continue
    p = sum(1 for k in (2, 4, 8, 12) if i >= k)
    hw = S >> p
    cin, cout = synth.CONV_LAYERS[i][1], synth.CONV_LAYERS[i][2]
    for grad in (0, 1):
        K, N = (cout, cin) if grad else (cin, cout)
        x = torch.rand(hw, hw, K, device="cuda") * 100
        y = torch.empty(hw, hw, N, device="cuda")
        lib.check(L.adpst_absmax(lib.ptr(x), x.numel(), lib.ptr(slot), lib.stream_ptr()))
        fn = L.adpst_vgg_conv_dgrad if grad else L.adpst_vgg_conv_forward
        ts = []
        for r in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.check(fn(ext.vgg._h, i, lib.ptr(x), hw, hw, lib.ptr(y), lib.ptr(slot), lib.stream_ptr()))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts[1:])
        f = 2.0 * hw * hw * 9 * cin * cout
        tot_f += f; tot_t += ms
        print("conv %2d %s %4dx%-4d K=%3d N=%3d: %.3f ms  %6.1f TFLOP/s" % (i, "dgrad" if grad else "fwd  ", hw, hw, K, N, ms, f / ms / 1e9), flush=True)
print("total %.3f ms, %.1f TFLOP/s" % (tot_t, tot_f / tot_t / 1e9))
