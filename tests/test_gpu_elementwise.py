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
