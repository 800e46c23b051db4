"""CPU: mask helpers against the reference's own outputs (golden), torch restatements against the numpy ones,
and self-consistency of the loss/Adam restatement."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import masks, matting, model


@pytest.mark.parametrize("tag", ["a", "b"])
def test_mask_helpers_match_reference(tag):
    g = golden("masks_%s.npz" % tag)
    d = masks.extract_segmentation_masks(g["seg"])
    keys = sorted(d)
    assert np.array_equal(np.array(keys, dtype=np.int64), g["keys"])       # class order: bit-exact
    m = np.stack(masks.mask_for_tf(d))
    assert m.dtype == np.float32 and np.array_equal(m, g["masks"])
    assert np.array_equal(masks.reduce_dict(d, (1,) + g["seg"].shape), g["reduced"])
    assert np.array_equal(g["reduced"].astype(np.uint8), g["seg"])          # round trip to the BGR image


def test_v2_torch_equals_numpy(synth):
    H, W = 10, 13
    img = synth.image(H, W, 21)[0].astype(np.float64)
    x = np.random.default_rng(22).random((H * W, 3))
    a = matting.V2Operator(img, 1e-7, 1)
    b = model.V2Torch(img, 1e-7, 1)
    np.testing.assert_allclose(b.means.numpy(), a.means, atol=1e-14)
    np.testing.assert_allclose(b.delta_inv.numpy(), a.delta_inv, rtol=1e-9)
    ya, yb = a.matmul(x), b.matmul(torch.as_tensor(x)).numpy()
    assert np.abs(ya - yb).max() < 1e-10 * np.abs(ya).max()


def test_photorealism_gradient_is_2Lx(synth):
    H, W = 8, 9
    img = synth.image(H, W, 23)
    lap = model.V2Torch(img[0], 1e-7, 1)
    x = torch.as_tensor(synth.image(H, W, 24), dtype=torch.float64).requires_grad_(True)
    loss = model.photorealism(x, lap)
    (g,) = torch.autograd.grad(loss, x)
    y = lap.matmul(x.detach().reshape(H * W, 3))
    assert torch.abs(g.reshape(H * W, 3) - 2 * y).max() < 1e-10 * torch.abs(y).max()


def test_resize_mask_half_pixel_centres():
    m = torch.zeros(1, 8, 8, 1)
    m[:, :, 3:] = 1.0
    r = model.resize_mask(m, (4, 4))[0, 0, :, 0]
    assert torch.allclose(r, torch.tensor([0.0, 0.5, 1.0, 1.0]))
    r = model.resize_mask(m, (2, 2))[0, 0, :, 0]     # scale 4: average of source px 1,2 / 5,6
    assert torch.allclose(r, torch.tensor([0.0, 1.0]))


def test_adam_tf_flavour_first_step():
    x = torch.tensor([0.5, 0.95, 0.02])
    g = torch.tensor([1.0, -2.0, 3.0])
    x1, m, v = model.adam_clip_step(x, g, torch.zeros(3), torch.zeros(3), 1)
    # t=1: alpha = lr*sqrt(1-b2)/(1-b1); m/(sqrt(v)+eps) = (1-b1) g / (sqrt(1-b2)|g| + eps) -> step ~ lr*sign(g)
    np.testing.assert_allclose(x1.numpy(), [0.4, 1.0, 0.0], atol=1e-6)


def test_small_train_state_runs_and_content_loss_zero_at_start(synth):
    H = W = 32
    w = synth.vgg_weights(seed=5)
    content = torch.as_tensor(synth.image(H, W, 0))
    style = torch.as_tensor(synth.image(H, W, 1))
    seg = synth.label_image(H, W, 2, 9, cell=16)
    ms = [torch.as_tensor(m) for m in masks.mask_for_tf(masks.extract_segmentation_masks(seg))]
    cfg = {"weights": {"content": 1.0, "style": 100.0, "nima": 0.0, "photo": 1e4},
           "matting_epsilon": 1e-7, "matting_window_radius": 1,
           "adam": {"lr": 0.1, "beta1": 0.9, "beta2": 0.999, "epsilon": 1e-8}}
    st = model.TrainState(content, style, w, cfg, ms, ms)
    d, g = st.loss_and_grad()
    assert d["Content loss"] == 0.0                    # transfer image starts as the content image
    assert d["Photorealism regualarization"] < 1e-3
    assert set(d) == {"Content loss", "Style loss", "NIMA loss", "Photorealism regualarization", "Total loss"}
    assert g.shape == (1, H, W, 3) and torch.isfinite(g).all()
    d2 = st.train_step()
    assert st.image.min() >= 0 and st.image.max() <= 1 and st.t == 1


def test_vgg_restatement_matches_torchvision_vgg19(synth):
    """The VGG19 arithmetic lives in Keras (un-vendored, un-pinned): the restatement is checked here against an INDEPENDENT
    implementation of the same published network, torchvision.models.vgg19 (features[0:30] = block1_conv1 .. relu5_1),
    loaded with the same seeded weights.  This pins topology, padding, pooling and the tap positions; it cannot pin
    TensorFlow's own kernels."""
    tv = pytest.importorskip("torchvision")
    weights = synth.vgg_weights(seed=11)
    net = tv.models.vgg19(weights=None).features[:30].double().eval()
    convs = [m for m in net if isinstance(m, torch.nn.Conv2d)]
    names = [it[0] for it in model.VGG_TOPOLOGY if it != "P"]
    assert len(convs) == len(names) == 13
    with torch.no_grad():
        for m, n in zip(convs, names):
            k, b = weights[n]
            m.weight.copy_(torch.as_tensor(k).double().permute(3, 2, 0, 1))      # HWIO -> OIHW
            m.bias.copy_(torch.as_tensor(b).double())
    img = torch.as_tensor(synth.image(40, 56, 12)).double()
    ours = model.vgg_forward(img, weights)
    x = (img * 255.0).flip(-1) - torch.tensor(model.CAFFE_MEAN_BGR, dtype=torch.float64)   # Keras caffe-mode preprocess
    x = x.permute(0, 3, 1, 2)
    taps = {1: "block1_conv1", 6: "block2_conv1", 11: "block3_conv1", 20: "block4_conv1", 22: "block4_conv2", 29: "block5_conv1"}
    with torch.no_grad():
        for i, layer in enumerate(net):
            x = layer(x)
            if i in taps:
                ref = x.permute(0, 2, 3, 1)
                got = ours[taps[i]]
                assert got.shape == ref.shape
                assert float((got - ref).abs().max()) <= 1e-9 * float(ref.abs().max()), taps[i]


def _loss_case(tag):
    g = golden("loss_%s.npz" % tag)
    t64 = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    ct = {k[3:]: t64(g[k]) for k in g.files if k.startswith("ct_")}
    co = {k[3:]: t64(g[k]) for k in g.files if k.startswith("co_")}
    st = {k[3:]: t64(g[k]) for k in g.files if k.startswith("st_")}
    so = {k[3:]: t64(g[k]) for k in g.files if k.startswith("so_")}
    K = int(g["K"])
    cm = [t64(m) for m in g["cmasks"]] if K else None
    sm = [t64(m) for m in g["smasks"]] if K else None
    return g, ct, co, st, so, cm, sm


@pytest.mark.parametrize("tag", ["a", "b"])
def test_loss_restatement_matches_reference_code(tag):
    """tests/golden/loss_*.npz come from the reference's OWN components/loss.py (Loss.__init__, compute_loss, iter_on_layers,
    the content / Gram / style / photorealism terms and the weighted total) executed over oracle/tf_shim.py."""
    g, ct, co, st, so, cm, sm = _loss_case(tag)
    image = torch.as_tensor(np.asarray(g["image"], np.float64))
    wp = float(g["photo_weight"])
    cfg = {"content": 1.0, "style": 100.0, "nima": 0.0, "photo": wp}
    lap = model.V2Torch(image[0].numpy(), 1e-7, 1) if wp > 0 else None
    d = model.compute_loss(image, {"content": co, "style": so}, ct, st, cfg, lap, cm, sm)
    assert abs(float(d["Content loss"]) - float(g["content_loss"])) <= 1e-12 * float(g["content_loss"])
    assert abs(float(d["Style loss"]) - float(g["style_loss"])) <= 1e-11 * float(g["style_loss"])
    if wp > 0:
        assert abs(float(d["Photorealism regualarization"]) - float(g["photo_loss"])) <= 1e-8 * float(g["photo_loss"])
    # (the photo term of case a is x^T L x at x = I: a 1e-5-sized remainder of O(1) terms, reproducible to ~1e-8 relative)
    assert abs(float(d["Total loss"]) - float(g["total_minus_nima"])) <= 1e-9 * float(g["total_minus_nima"])
    # dict order of the reference: content, style, nima, [photo], total (loss.py:72-76)
    assert [k for k in d.keys()] == [str(k) for k in g["keys"]]
