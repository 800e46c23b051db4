"""GPU parity of the VGG19 extractor, masked Gram / style / content terms, the total loss, its gradient and the full
train_step against the torch-CPU float64 oracle on identical synthetic inputs (random-init VGG19, synthetic masks).

Tolerance: 1e-5 relative (north star) for every loss scalar and, in max-norm, for feature maps and gradients.
Measured float32 noise floor of the restatement itself (torch-CPU float32 vs float64): ~1e-7.
"""
import argparse
import importlib

import numpy as np
import pytest
import torch

from conftest import PKG_NAME
from oracle import masks as omasks
from oracle import model
from oracle import parity

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _m(name):
    return importlib.import_module(PKG_NAME + "." + name)


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _assert_gradient_close(g, go, ext, x, weights, synth):
    """Max-norm 1e-5 everywhere outside the receptive fields of ReLU units whose decision differs between the float32 GPU
    forward pass and the float64 oracle (oracle/parity.py): a pre-activation within rounding distance of zero flips its
    mask in ANY float32 implementation and changes the gradient inside that one unit's receptive field only."""
    ref = model.vgg_forward(x.cpu().double(), weights)
    ref_acts = [ref[name] for name, _, _ in synth.CONV_LAYERS]
    return parity.assert_gradient_close(g, go, ext.last.acts, ref_acts, TOL)


def _args(**kw):
    d = dict(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=1e4,
             matting_epsilon=1e-7, matting_window_radius=1, adam_lr=0.1, adam_beta1=0.9, adam_beta2=0.999,
             adam_epsilon=1e-8, iter=3)
    d.update(kw)
    return argparse.Namespace(**d)


def _cfg(a):
    return {"weights": {"content": a.content_weight, "style": a.style_weight, "nima": 0.0, "photo": a.regularization_weight,
                        "tv": getattr(a, "tv_weight", 0.0)},
            "matting_epsilon": a.matting_epsilon, "matting_window_radius": a.matting_window_radius,
            "adam": {"lr": a.adam_lr, "beta1": a.adam_beta1, "beta2": a.adam_beta2, "epsilon": a.adam_epsilon}}


@pytest.fixture(scope="module")
def weights(synth):
    return synth.vgg_weights(seed=5)


@pytest.mark.parametrize("H,W", [(64, 64), (37, 50), (16, 16), (96, 138), (144, 72)])     # 144x72: 16x8 pixel tiles (narrow maps)
def test_vgg_forward_all_taps(H, W, weights, synth):
    vgg = _m("components.VGG19.model")
    names = [n for n, _, _ in synth.CONV_LAYERS]
    ext = vgg.StyleContentModel(names[:1], names[1:], weights=weights)
    img = synth.image(H, W, 0)
    out = ext(torch.as_tensor(img).cuda())
    ref = model.vgg_forward(torch.as_tensor(img), weights)
    got = dict(out["content"]); got.update(out["style"])
    for n in names:
        assert tuple(got[n].shape) == tuple(ref[n].shape), n
        assert _rel(got[n].cpu().numpy(), ref[n].numpy()) < TOL, n


@pytest.mark.parametrize("H,W", [(32, 32), (37, 50), (90, 40)])
def test_vgg_backward_matches_autograd(H, W, weights, synth):
    """d(sum_i <seed_i, layer_i>)/d(image) with random seeds on the six tapped layers."""
    vgg = _m("components.VGG19.model")
    ext = vgg.StyleContentModel(model.CONTENT_LAYERS, model.STYLE_LAYERS, weights=weights)
    img = synth.image(H, W, 3)
    out = ext(torch.as_tensor(img).cuda())
    rng = np.random.default_rng(4)
    seeds, flat = {}, dict(out["content"]); flat.update(out["style"])
    for n, t in flat.items():
        seeds[n] = rng.standard_normal(tuple(t.shape)).astype(np.float32)
    g = ext.backward({n: torch.as_tensor(s).cuda() for n, s in seeds.items()})
    x = torch.as_tensor(img, dtype=torch.float64).requires_grad_(True)
    ref = model.vgg_forward(x, weights)
    tot = sum((ref[n] * torch.as_tensor(s, dtype=torch.float64)).sum() for n, s in seeds.items())
    (gr,) = torch.autograd.grad(tot, x)
    assert _rel(g.cpu().numpy(), gr.numpy()) < TOL


@pytest.mark.parametrize("H,W", [(32, 32), (37, 51)])
def test_vgg_backward_with_seeds_on_every_layer(H, W, weights, synth):
    """Seeds on all 13 convolutions: the layers in front of a max-pool then carry a seed as well (the un-pooling kernel adds it
    before the ReLU mask; the six tapped layers of the real objective leave those layers without one)."""
    vgg = _m("components.VGG19.model")
    names = [n for n, _, _ in synth.CONV_LAYERS]
    ext = vgg.StyleContentModel(names[:1], names[1:], weights=weights)
    img = synth.image(H, W, 5)
    out = ext(torch.as_tensor(img).cuda())
    rng = np.random.default_rng(6)
    flat = dict(out["content"]); flat.update(out["style"])
    seeds = {n: rng.standard_normal(tuple(t.shape)).astype(np.float32) for n, t in flat.items()}
    g = ext.backward({n: torch.as_tensor(s).cuda() for n, s in seeds.items()})
    x = torch.as_tensor(img, dtype=torch.float64).requires_grad_(True)
    ref = model.vgg_forward(x, weights)
    tot = sum((ref[n] * torch.as_tensor(s, dtype=torch.float64)).sum() for n, s in seeds.items())
    (gr,) = torch.autograd.grad(tot, x)
    assert _rel(g.cpu().numpy(), gr.numpy()) < TOL


@pytest.mark.parametrize("path", ["tensor", "simt"])
@pytest.mark.parametrize("hw,C,K", [((64, 64), 64, 3), ((25, 40), 128, 1), ((21, 37), 256, 4), ((16, 16), 512, 2),
                                    ((40, 48), 128, 8)])
def test_masked_gram_and_style_gradient(hw, C, K, path):
    k = _m("kernels")
    HW = hw[0] * hw[1]
    rng = np.random.default_rng(HW + C)
    F = (rng.random((HW, C)) * 50).astype(np.float32)
    S = (rng.random((HW // 2 + 3, C)) * 50).astype(np.float32)
    if K > 1:
        lab = rng.integers(0, K, HW)
        if K == 8:                                             # blocky labels: some classes absent from whole tiles
            lab = np.kron(rng.integers(0, K, (hw[0] // 8, hw[1] // 8)), np.ones((8, 8), dtype=np.int64)).reshape(-1)
        m = np.stack([(lab == i).astype(np.float32) for i in range(K)])
        soft = rng.random((K, HW)) < 0.05                       # some fractional boundary pixels
        m = np.where(soft, rng.random((K, HW)).astype(np.float32), m).astype(np.float32)
        ms = rng.random((K, S.shape[0])).astype(np.float32)
    else:
        m, ms = None, None
    Fd = torch.as_tensor(F).cuda()
    md = None if m is None else torch.as_tensor(m).cuda()
    F3 = Fd.reshape(hw[0], hw[1], C)
    G = k.gram_masked(F3, md, K, path=path, patches=k.gram_patch_lists(md, hw[0], hw[1], K, "cuda") if path == "tensor" else None)
    A = k.gram_masked(torch.as_tensor(S).cuda(), None if ms is None else torch.as_tensor(ms).cuda(), K, path="simt")
    Ft = torch.as_tensor(F, dtype=torch.float64).requires_grad_(True)
    St = torch.as_tensor(S, dtype=torch.float64)
    loss = 0.0
    for i in range(K):
        mi = torch.ones(HW, dtype=torch.float64) if m is None else torch.as_tensor(m[i], dtype=torch.float64)
        si = torch.ones(S.shape[0], dtype=torch.float64) if ms is None else torch.as_tensor(ms[i], dtype=torch.float64)
        g_t = model.gram_matrix(Ft.reshape(1, 1, HW, C), mi)
        g_s = model.gram_matrix(St.reshape(1, 1, -1, C), si)
        assert _rel(G[i].cpu().numpy(), g_t.detach().numpy()) < TOL
        assert _rel(A[i].cpu().numpy(), g_s.numpy()) < TOL
        loss = loss + torch.mean((g_s - g_t) ** 2) / (2 * float(C) ** 2 * float(HW) ** 2)
    (gF,) = torch.autograd.grad(loss * 50.0, Ft)
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    dF = torch.empty_like(Fd)
    k.style_layer_backward(Fd.reshape(hw[0], hw[1], C), md, K, G, A, 0.5, 50.0, acc, dF, path=path)
    assert abs(float(acc) - 0.5 * float(loss.detach())) < TOL * 0.5 * float(loss.detach())
    assert _rel(dF.cpu().numpy(), gF.numpy()) < 5 * TOL     # G - A cancels digits: allow for float32 Gram rounding
    dF2 = dF.clone()
    k.style_layer_backward(Fd.reshape(hw[0], hw[1], C), md, K, G, A, 0.5, 50.0, acc, dF2, accumulate=True, path=path)
    assert _rel(dF2.cpu().numpy(), 2 * gF.numpy()) < 5 * TOL


def test_masks_above_one_and_wide_dynamic_range():
    """The tensor-core kernels scale their FP16 operands by max|F| (and max|mask|): masks above 1 and feature maps with a
    huge dynamic range (values 2^-30 of the maximum) must neither overflow nor lose the max-norm accuracy."""
    k = _m("kernels")
    h, w, C, K = 24, 32, 128, 3
    rng = np.random.default_rng(5)
    F = (rng.random((h * w, C)) * 2000).astype(np.float32)
    F[::7] *= 2.0 ** -30                                   # rows far below the FP16 normal range after scaling
    lab = rng.integers(0, K, h * w)
    m = (np.stack([(lab == i) for i in range(K)]).astype(np.float32) * 3.0).astype(np.float32)      # weights up to 9
    Fd, md = torch.as_tensor(F).cuda(), torch.as_tensor(m).cuda()
    F3 = Fd.reshape(h, w, C)
    Gt = k.gram_masked(F3, md, K, path="tensor", patches=k.gram_patch_lists(md, h, w, K, "cuda"))
    X = torch.as_tensor(F, dtype=torch.float64)
    ref = torch.stack([(X * float(1)) .T @ (X * torch.as_tensor(m[i], dtype=torch.float64)[:, None] ** 2) for i in range(K)])
    assert torch.isfinite(Gt).all()
    assert _rel(Gt.cpu().numpy(), ref.numpy()) < TOL
    A = torch.zeros_like(Gt)
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    dT, dS = torch.empty_like(Fd), torch.empty_like(Fd)
    k.style_layer_backward(F3, md, K, Gt, A, 1.0, 1.0, acc, dT, path="tensor")
    k.style_layer_backward(F3, md, K, Gt, A, 1.0, 1.0, acc, dS, path="simt")
    assert torch.isfinite(dT).all()
    assert _rel(dT.cpu().numpy(), dS.cpu().numpy()) < 5 * TOL


def test_absmax_slot():
    k = _m("kernels")
    for n in (1, 3, 4, 1000, 4099):
        x = (torch.rand(n, device="cuda") - 0.5) * 37.0
        slot = k.absmax_slot(x)
        assert float(slot.view(torch.float32)[0]) == float(x.abs().max())
    assert int(k.absmax_slot(torch.zeros(16, device="cuda"))[0]) == 0


def test_repeated_runs_are_stable(weights, synth):
    """Two independent set-ups run the same 25 eager steps: no launch failure (pipeline races show up as rare ones) and the
    same losses (reductions are order-fixed except the float64 atomics of the style term)."""
    st = _m("style_transfer")
    args = _args()
    hist = []
    for run in range(2):
        ext, loss, _, c_dev = _setup(96, 128, 3, weights, synth, args)
        step = st.make_train_step(ext, loss, st.Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon))
        x = c_dev.clone()
        out = []
        for _ in range(25):
            d = step(x)
            out.append(float(d["Total loss"]))
        torch.cuda.synchronize()
        hist.append(out)
    for a, b in zip(*hist):
        assert abs(a - b) <= 1e-6 * abs(a)


def _setup(H, W, K, weights, synth, args, matting="v2", Hs=None, Ws=None):
    vgg, lossm = _m("components.VGG19.model"), _m("components.loss")
    Hs, Ws = Hs or H, Ws or W
    content, style = synth.image(H, W, 0), synth.image(Hs, Ws, 1)
    cm = sm = cm_o = sm_o = None
    if K:
        cell = max(4, min(H, W, Hs, Ws) // 4)
        segc, segs = synth.label_image(H, W, K, 9, cell=cell), synth.label_image(Hs, Ws, K, 10, cell=cell)
        sem = _m("components.semantic_merge")
        cm, sm = sem.mask_for_tf(sem.extract_segmentation_masks(segc)), sem.mask_for_tf(sem.extract_segmentation_masks(segs))
        cm_o = [torch.as_tensor(m) for m in omasks.mask_for_tf(omasks.extract_segmentation_masks(segc))]
        sm_o = [torch.as_tensor(m) for m in omasks.mask_for_tf(omasks.extract_segmentation_masks(segs))]
        assert len(cm) == len(cm_o) == K and all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(cm, cm_o))
    ext = vgg.StyleContentModel(model.CONTENT_LAYERS, model.STYLE_LAYERS, shape=(None, None, 3), weights=weights)
    c_dev, s_dev = torch.as_tensor(content).cuda(), torch.as_tensor(style).cuda()
    loss = lossm.Loss(ext(c_dev)["content"], ext(s_dev)["style"], args, cm, sm, matting=matting)
    if args.regularization_weight > 0:
        loss.initialize_matting_laplacian(c_dev[0].double())
    ora = model.TrainState(torch.as_tensor(content), torch.as_tensor(style), weights, _cfg(args), cm_o, sm_o)
    return ext, loss, ora, c_dev


@pytest.mark.parametrize("H,W,K,photo", [(64, 64, 3, 1e4), (64, 64, 0, 0.0), (37, 50, 2, 1e4), (48, 80, 4, 0.0)])
def test_total_loss_and_image_gradient(H, W, K, photo, weights, synth):
    args = _args(regularization_weight=photo)
    ext, loss, ora, c_dev = _setup(H, W, K, weights, synth, args)
    # evaluate away from the content image (at x = content the content term and its gradient vanish)
    pert = np.sign(synth.image(H, W, 3) - 0.5).astype(np.float32) * 0.1
    x = torch.clamp(c_dev + torch.as_tensor(pert).cuda(), 0, 1).contiguous()
    d = loss(x, ext(x, reuse=True))
    g = loss.gradient(ext)
    do, go = ora.loss_and_grad(x.cpu().double())
    assert set(d) == set(do) or (photo == 0 and set(d) == set(do) - {"Photorealism regualarization"}) or set(d) == set(do)
    for name, v in d.items():
        assert abs(float(v) - do[name]) <= TOL * abs(do[name]) + 1e-12, name
    _assert_gradient_close(g.cpu().numpy(), go.numpy(), ext, x, weights, synth)


@pytest.mark.parametrize("H,W,K,photo,tv", [(64, 64, 3, 1e4, 5.0), (37, 50, 0, 0.0, 2.0)])
def test_tv_extension_total_and_gradient(H, W, K, photo, tv, weights, synth):
    """tv_weight > 0 (extension, tf.image.total_variation semantics): extra dictionary key before 'Total loss', total and
    gradient include the term; with tv_weight = 0 the dictionary is the reference's (checked by every other test)."""
    args = _args(regularization_weight=photo, tv_weight=tv)
    ext, loss, ora, c_dev = _setup(H, W, K, weights, synth, args)
    pert = np.sign(synth.image(H, W, 3) - 0.5).astype(np.float32) * 0.1
    x = torch.clamp(c_dev + torch.as_tensor(pert).cuda(), 0, 1).contiguous()
    d = loss(x, ext(x, reuse=True))
    g = loss.gradient(ext)
    do, go = ora.loss_and_grad(x.cpu().double())
    assert list(d) == list(do) and list(d)[-2:] == ["Total variation loss", "Total loss"]
    for name, v in d.items():
        assert abs(float(v) - do[name]) <= TOL * abs(do[name]) + 1e-12, name
    _assert_gradient_close(g.cpu().numpy(), go.numpy(), ext, x, weights, synth)


def test_style_and_content_sizes_differ(weights, synth):
    """Defaults of the reference: content 96x138, style 96x168 (blanc.jpg / bear.jpeg)."""
    args = _args(regularization_weight=0.0)
    ext, loss, ora, c_dev = _setup(48, 69, 2, weights, synth, args, Hs=48, Ws=84)
    x = torch.clamp(c_dev * 0.9 + 0.05, 0, 1).contiguous()
    d = loss(x, ext(x, reuse=True))
    g = loss.gradient(ext)
    do, go = ora.loss_and_grad(x.cpu().double())
    for name, v in d.items():
        assert abs(float(v) - do[name]) <= TOL * abs(do[name]) + 1e-12, name
    _assert_gradient_close(g.cpu().numpy(), go.numpy(), ext, x, weights, synth)


@pytest.mark.parametrize("graph", [False, True])
def test_train_steps_follow_oracle(graph, weights, synth):
    st = _m("style_transfer")
    args = _args()
    ext, loss, ora, c_dev = _setup(64, 64, 3, weights, synth, args)
    opt = st.Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon)
    step = st.make_train_step(ext, loss, opt, use_cuda_graph=graph)
    x = c_dev.clone()
    for it in range(3):
        d = {k: float(v) for k, v in step(x).items()}
        do = ora.train_step()
        for name in d:
            assert abs(d[name] - do[name]) <= 1e-4 * abs(do[name]) + 1e-6, (it, name)
        # Adam's first steps are +-lr*sign(g): a pixel whose gradient is ~0 may flip; count, don't max
        diff = (x.cpu().double() - ora.image).abs()
        assert float((diff > 1e-3).double().mean()) < 1e-3, it
    assert opt.iterations == 3
    mse = float(((x.cpu().double() - ora.image) ** 2).mean())
    assert 10 * np.log10(1.0 / max(mse, 1e-30)) > 50.0           # PSNR bar of the north star


def test_hundred_adam_iterations_psnr(weights, synth):
    """North star: after 100 Adam iterations the image is within 50 dB PSNR of the reference (here: the float64 oracle)."""
    st = _m("style_transfer")
    args = _args()
    ext, loss, ora, c_dev = _setup(64, 64, 3, weights, synth, args)
    opt = st.Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon)
    step = st.make_train_step(ext, loss, opt, use_cuda_graph=True)
    x = c_dev.clone()
    for it in range(100):
        d = step(x)
        do = ora.train_step()
    assert opt.iterations == 100
    mse = float(((x.cpu().double() - ora.image) ** 2).mean())
    psnr = 10 * np.log10(1.0 / max(mse, 1e-30))
    rel = abs(float(d["Total loss"]) - do["Total loss"]) / abs(do["Total loss"])
    print("PSNR after 100 iterations: %.1f dB, total loss rel. diff %.2e" % (psnr, rel))
    assert psnr > 50.0
    assert rel < 1e-3


@pytest.mark.parametrize("tag", ["a", "b"])
def test_loss_matches_reference_code_golden(tag):
    """Against tests/golden/loss_*.npz: the loss dict of the reference's own components/loss.py (oracle/make_golden.py)
    for given feature maps, masks and image -- content, masked-Gram style, photorealism, weighted total, key order."""
    from conftest import golden
    lossm = _m("components.loss")
    g = golden("loss_%s.npz" % tag)
    dev = lambda a: torch.as_tensor(np.asarray(a, np.float32)).cuda().contiguous()
    ct = {k[3:]: dev(g[k]) for k in g.files if k.startswith("ct_")}
    co = {k[3:]: dev(g[k]) for k in g.files if k.startswith("co_")}
    st_ = {k[3:]: dev(g[k]) for k in g.files if k.startswith("st_")}
    so = {k[3:]: dev(g[k]) for k in g.files if k.startswith("so_")}
    K, wp = int(g["K"]), float(g["photo_weight"])
    cm = [np.asarray(m, np.float32) for m in g["cmasks"]] if K else None
    sm = [np.asarray(m, np.float32) for m in g["smasks"]] if K else None
    image = dev(g["image"])
    loss = lossm.Loss(ct, st_, _args(regularization_weight=wp), cm, sm)
    if wp > 0:
        loss.initialize_matting_laplacian(image[0].double())
    d = loss(image, {"content": co, "style": so})
    assert abs(float(d["Content loss"]) - float(g["content_loss"])) <= TOL * float(g["content_loss"])
    assert abs(float(d["Style loss"]) - float(g["style_loss"])) <= TOL * float(g["style_loss"])
    if wp > 0:
        assert abs(float(d["Photorealism regualarization"]) - float(g["photo_loss"])) <= 1e-5 * float(g["photo_loss"])
    assert abs(float(d["Total loss"]) - float(g["total_minus_nima"])) <= TOL * float(g["total_minus_nima"])
    assert list(d.keys()) == [str(k) for k in g["keys"]]


def test_nima_weight_is_refused(weights, synth):
    lossm = _m("components.loss")
    t = {"block4_conv2": torch.zeros(1, 2, 2, 512, device="cuda")}
    with pytest.raises(NotImplementedError):
        lossm.Loss(t, {}, _args(nima_weight=1e5))


def test_script_main_with_segmentation_files(tmp_path, synth, monkeypatch, capsys):
    """python style_transfer.py end to end (reference :207-384): *_seg.png mask files -> extract_segmentation_masks ->
    mask_for_tf, meta.json, loss printing, per-iteration scalars, intermediate and best image; and the same run through
    the classes directly gives the same losses."""
    import json
    import cv2
    st = _m("style_transfer")
    sem = _m("components.semantic_merge")
    H, W, K = 48, 64, 3
    monkeypatch.chdir(tmp_path)
    content, style = synth.smooth_image(H, W, 11), synth.smooth_image(H, W, 12)
    cv2.imwrite("content.png", (content[0] * 255).round().astype(np.uint8)[:, :, ::-1])
    cv2.imwrite("style.png", (style[0] * 255).round().astype(np.uint8)[:, :, ::-1])
    (tmp_path / "raw_seg").mkdir()
    segc, segs = synth.label_image(H, W, K, 9, cell=16), synth.label_image(H, W, K, 10, cell=16)
    cv2.imwrite("raw_seg/content_seg.png", segc)
    cv2.imwrite("raw_seg/style_seg.png", segs)
    w = synth.vgg_weights(seed=5)
    np.savez("vgg.npz", **{n + "/kernel": k for n, (k, b) in w.items()}, **{n + "/bias": b for n, (k, b) in w.items()})
    argv = ["-c", "content.png", "-s", "style.png", "-o", "out.png", "--iter", "6", "--matting_window_radius", "1",
            "--matting_epsilon", "1e-7", "--vgg_weights", "vgg.npz", "--experiment_name", "t", "--use_masks",
            "--print_loss_interval", "2", "--intermediate_result_interval", "3"]
    best, hist = st.main(argv)
    out = capsys.readouterr().out
    assert "Load segmentation from files." in out and out.count("[Iter ") == 3 and "Average time per epoch" in out
    exp = tmp_path / "experiments" / "t"
    meta = json.loads((exp / "meta.json").read_text())
    assert meta["load_segmentation"] is True and meta["iter"] == 6 and meta["content"] == "content.png"
    assert sorted(p.name for p in (exp / "iter").iterdir()) == ["iter_3.png", "iter_6.png"]
    lines = [json.loads(l) for l in (tmp_path / "logs" / "t" / "scalars.jsonl").read_text().splitlines()]
    assert [l["step"] for l in lines] == list(range(1, 7)) and "Photorealism regualarization" in lines[0]
    assert len(hist) == 6 and tuple(best.shape) == (1, H, W, 3)
    saved = cv2.imread(str(exp / "out.png"))[:, :, ::-1]
    assert np.array_equal(saved, st.tensor_to_image(best).cpu().numpy())
    # the *_seg.png files are rewritten from the extracted masks (reference :261-264) without changing them
    assert np.array_equal(cv2.imread("raw_seg/content_seg.png"), segc)
    # same run through the classes, masks straight from the label images: identical first-iteration losses
    a = st.build_parser().parse_args(argv)
    cm = sem.mask_for_tf(sem.extract_segmentation_masks(segc)); sm = sem.mask_for_tf(sem.extract_segmentation_masks(segs))
    _, hist2 = st.style_transfer(st.load_image("content.png"), st.load_image("style.png"), a, cm, sm, vgg_weights="vgg.npz")
    for name, v in hist[0].items():
        assert abs(v - hist2[0][name]) <= 1e-6 * abs(v) + 1e-12, name
    # without --use_masks the script behaves like the reference as shipped (masks commented out, :308-309): K = 1
    _, hist3 = st.main([x for x in argv if x != "--use_masks"])
    assert abs(hist3[0]["Style loss"] - hist[0]["Style loss"]) > 1e-6 * abs(hist[0]["Style loss"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process(weights, synth):
    """Per-device kernel set-up (cudaFuncSetAttribute, SM count) and launch-device guards: one train step on cuda:0, then the
    same pair on cuda:1 in the same process while cuda:0 stays the current device; identical losses."""
    st = _m("style_transfer")
    vgg, lossm, sem = _m("components.VGG19.model"), _m("components.loss"), _m("components.semantic_merge")
    args = _args()
    H, W, K = 64, 96, 3
    content, style = synth.image(H, W, 0), synth.image(H, W, 1)
    cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, K, 9, cell=16)))
    sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, K, 10, cell=16)))
    results = []
    for dev in ("cuda:0", "cuda:1"):
        ext = vgg.StyleContentModel(model.CONTENT_LAYERS, model.STYLE_LAYERS, weights=weights, device=dev)
        c_dev, s_dev = torch.as_tensor(content).to(dev), torch.as_tensor(style).to(dev)
        loss = lossm.Loss(ext(c_dev)["content"], ext(s_dev)["style"], args, cm, sm)
        loss.initialize_matting_laplacian(c_dev[0].double())
        step = st.make_train_step(ext, loss, st.Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon))
        x = c_dev.clone()
        for _ in range(2):
            d = {k: float(v) for k, v in step(x).items()}
        torch.cuda.synchronize(dev)
        results.append((d, x.cpu()))
    assert torch.cuda.current_device() == 0
    for name, v in results[0][0].items():
        assert abs(v - results[1][0][name]) <= 1e-6 * abs(v) + 1e-12, name
    assert float((results[0][1] - results[1][1]).abs().max()) < 1e-5
