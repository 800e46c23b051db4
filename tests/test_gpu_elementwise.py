"""GPU parity: fused Adam+clip, content MSE (+gradient), mask resize -- against the torch-CPU oracle."""
import importlib

import numpy as np
import pytest
import torch

from conftest import PKG_NAME
from oracle import model

pytestmark = pytest.mark.gpu


def _k():
    return importlib.import_module(PKG_NAME + ".kernels")


@pytest.mark.parametrize("n", [3, 5 * 7 * 3, 64 * 64 * 3, 1000003])
def test_adam_clip_three_steps(n):
    k = _k()
    rng = np.random.default_rng(n)
    x0 = rng.random(n).astype(np.float32)
    x = torch.as_tensor(x0).cuda()
    st = k.AdamState(x)
    xo = torch.as_tensor(x0.astype(np.float64))
    mo, vo = torch.zeros_like(xo), torch.zeros_like(xo)
    gmax = 0.0
    for t in range(1, 4):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3)).astype(np.float32)
        gmax = max(gmax, float(np.abs(g).max()))
        k.adam_clip_step(x, torch.as_tensor(g).cuda(), st, 0.1, 0.9, 0.999, 1e-8)
        xo, mo, vo = model.adam_clip_step(xo, torch.as_tensor(g.astype(np.float64)), mo, vo, t)
        assert st.step == t
        np.testing.assert_allclose(x.cpu().numpy(), xo.numpy(), atol=2e-6, rtol=0)     # float32 state vs float64 oracle
        # float32 slots: m = b1*m + (1-b1)*g cancels when g changes sign, so the error is absolute in |g|
        np.testing.assert_allclose(st.m.cpu().numpy(), mo.numpy(), rtol=1e-5, atol=1e-6 * gmax)
        np.testing.assert_allclose(st.v.cpu().numpy(), vo.numpy(), rtol=1e-5, atol=1e-6 * gmax * gmax)
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0


@pytest.mark.parametrize("src,dst", [((64, 64), (32, 32)), ((64, 64), (16, 16)), ((64, 64), (4, 4)),
                                     ((96, 138), (48, 69)), ((96, 138), (24, 34)), ((33, 21), (33, 21)),
                                     ((17, 9), (8, 4))])
def test_resize_bilinear_half_pixel(src, dst):
    k = _k()
    m = (np.random.default_rng(1).random(src) > 0.5).astype(np.float32)
    got = k.resize_bilinear(torch.as_tensor(m).cuda(), dst).cpu().numpy()
    want = model.resize_mask(torch.as_tensor(m.astype(np.float64))[None, :, :, None], dst)[0, :, :, 0].numpy()
    np.testing.assert_allclose(got, want, atol=1e-6)


def test_content_layer_value_and_gradient():
    k = _k()
    rng = np.random.default_rng(3)
    t = rng.standard_normal((1, 16, 16, 512)).astype(np.float32) * 50
    o = rng.standard_normal((1, 16, 16, 512)).astype(np.float32) * 50
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    d = torch.empty(t.shape, dtype=torch.float32, device="cuda")
    k.content_layer(torch.as_tensor(t).cuda(), torch.as_tensor(o).cuda(), 0.5, 0.5, acc, d)
    oo = torch.as_tensor(o.astype(np.float64)).requires_grad_(True)
    loss = 0.5 * model.layer_content_loss(torch.as_tensor(t.astype(np.float64)), oo)
    (g,) = torch.autograd.grad(loss, oo)
    assert abs(float(acc) - float(loss.detach())) < 1e-9 * float(loss.detach())
    assert np.abs(d.cpu().numpy() - g.numpy()).max() < 1e-6 * np.abs(g.numpy()).max()
    k.content_layer(torch.as_tensor(t).cuda(), torch.as_tensor(o).cuda(), 0.5, 0.5, acc, d, accumulate=True)
    assert np.abs(d.cpu().numpy() - 2 * g.numpy()).max() < 2e-6 * np.abs(g.numpy()).max()
    assert abs(float(acc) - 2 * float(loss.detach())) < 1e-9 * float(loss.detach())


@pytest.mark.parametrize("H,W", [(1, 1), (1, 7), (5, 1), (37, 50), (64, 64), (130, 257)])
def test_total_variation_extension(H, W):
    """EXTENSION (not in the reference): value and gradient of tf.image.total_variation against the torch-CPU float64
    oracle, on an image quantised to k/8 so that many differences are exactly zero (sgn(0) = 0)."""
    from oracle import model
    k = importlib.import_module(PKG_NAME + ".kernels")
    rng = np.random.default_rng(H * 1000 + W)
    img = (np.round(rng.random((1, H, W, 3)) * 8) / 8).astype(np.float32)
    x = torch.as_tensor(img).cuda()
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    g = torch.full_like(x, float("nan"))
    k.tv_loss(x, 0.5, 3.0, acc, g)
    xo = torch.as_tensor(img, dtype=torch.float64).requires_grad_(True)
    tv = model.total_variation(xo)
    (go,) = torch.autograd.grad(3.0 * tv, xo) if H * W > 1 else (torch.zeros_like(xo),)
    assert abs(float(acc) - 0.5 * float(tv)) <= 1e-12 * max(float(tv), 1.0)
    assert torch.equal(g.cpu().double(), go)                       # sums of +-3: exact
    g2 = g.clone()
    k.tv_loss(x, 0.5, 3.0, None, g2, accumulate=True)
    assert torch.equal(g2.cpu().double(), 2 * go)
    # scalar restricted to a column window (spatially tiled runs): the windows partition the sum
    if W >= 4:
        parts = []
        for lo, hi in ((0, W // 2), (W // 2, W)):
            a = torch.zeros(1, dtype=torch.float64, device="cuda")
            k.tv_loss(x, 1.0, 0.0, a, None, own_cols=(lo, hi))
            parts.append(float(a))
        assert abs(sum(parts) - float(tv)) <= 1e-12 * max(float(tv), 1.0)
