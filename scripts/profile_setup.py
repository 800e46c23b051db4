"""Where does the per-pair set-up time go (BASELINE configs[2])?  cProfile of bench.build_pair + graph capture at 512x512."""
import cProfile, importlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
st, vggm, synth = bench.mod("style_transfer"), bench.mod("components.VGG19.model"), bench.mod("synth")
S, K = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 4
hp = bench.hyper(1.0)
ext = vggm.StyleContentModel(st.CONTENT_LAYERS, st.STYLE_LAYERS, shape=(None, None, 3), weights=synth.vgg_weights())
def one(p):
    loss, opt, content, _ = bench.build_pair(ext, S, K, hp, (2 * p, 2 * p + 1, 100 + p, 200 + p))
    x = content.clone()
    step = st.make_train_step(ext, loss, opt, use_cuda_graph=True)
    step(x); torch.cuda.synchronize()
    return step
one(0); one(1)
t0 = time.perf_counter(); one(2); print("one pair set-up: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable(); one(3); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
