"""GPU: one image split into column strips (spatial tiling, BASELINE configs[3]) must reproduce the single-device
optimisation: same loss values, same updated image.  The ranks are emulated inside one process (one thread per rank, the
halo exchanges and the Gram reduction go through tiled.ThreadComm); the NCCL back end runs the same step() and is checked
against the single-device run inside bench.py (tiled_4k.parity_vs_single_device) on >= 2 GPUs."""
import argparse
import importlib

import numpy as np
import pytest
import torch

from conftest import PKG_NAME

pytestmark = pytest.mark.gpu


def _m(name):
    return importlib.import_module(PKG_NAME + "." + name)


def _args():
    return argparse.Namespace(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=1e4,
                              matting_epsilon=1e-7, matting_window_radius=1, adam_lr=0.1, adam_beta1=0.9, adam_beta2=0.999,
                              adam_epsilon=1e-8)


@pytest.mark.parametrize("world,W,K,halo", [(2, 512, 3, None), (4, 640, 0, None), (2, 512, 3, "peer"), (4, 640, 2, "peer")])
def test_tiled_matches_single_device(world, W, K, halo, synth):
    st, tiled, vgg, lossm, sem = _m("style_transfer"), _m("tiled"), _m("components.VGG19.model"), _m("components.loss"), \
        _m("components.semantic_merge")
    H = 48
    args = _args()
    weights = synth.vgg_weights(seed=5)
    content, style = synth.image(H, W, 0), synth.image(H, W, 1)
    cm = sm = None
    if K:
        cm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, K, 9, cell=16)))
        sm = sem.mask_for_tf(sem.extract_segmentation_masks(synth.label_image(H, W, K, 10, cell=16)))
    # single-device reference run (same kernels, whole image)
    ext = vgg.StyleContentModel(st.CONTENT_LAYERS, st.STYLE_LAYERS, weights=weights)
    c_dev, s_dev = torch.as_tensor(content).cuda(), torch.as_tensor(style).cuda()
    loss = lossm.Loss(ext(c_dev)["content"], ext(s_dev)["style"], args, cm, sm)
    loss.initialize_matting_laplacian(c_dev[0].double())
    step = st.make_train_step(ext, loss, st.Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon))
    x = c_dev.clone()
    ref = [{k: float(v) for k, v in step(x).items()} for _ in range(3)]
    # tiled run, ranks emulated in this process
    ranks = tiled.make_emulated(content, style, args, cm, sm, weights, world, halo=halo)
    assert all((r.halo is not None) == (halo == "peer") for r in ranks)
    got = tiled.run_emulated(ranks, 3)
    assert ranks[0].exchange_bytes()["exchanges"] == 11 and all(r.tile.halos == tiled.LEVEL_HALO == (4, 4, 8, 4, 2) for r in ranks)
    # Iteration 0 evaluates identical images: 2e-5.  Afterwards the images themselves may differ in a few pixels: Adam's first
    # steps are +-lr * sign(g), so a last-bit difference in a gradient that is ~0 moves that pixel by 2 lr (the image check
    # below counts such flips); the loss values of later iterations are therefore only comparable to ~1e-4.
    for it in range(3):
        tol = 2e-5 if it == 0 else 5e-4
        for name, v in ref[it].items():
            assert abs(got[it][name] - v) <= tol * abs(v) + 1e-9, (it, name, got[it][name], v)
    stitched = torch.cat([r.own_strip() for r in ranks], dim=1)
    diff = (stitched - x[0]).abs()
    assert float((diff > 1e-3).float().mean()) < 1e-3          # Adam steps of +-lr: count flips, do not take the max
    # every rank's halo holds the neighbours' updated pixels
    for r in ranks:
        t = r.tile
        assert float((r.image[0] - x[0, :, t.ext_lo:t.ext_hi]).abs().gt(1e-3).float().mean()) < 1e-3


def test_peer_halo_kernels():
    """adpst_halo_push / adpst_halo_pull on three mailboxes of one process: every halo column ends up holding the
    neighbour's own columns, the scale word is raised to max|received|, and the same (slot, offset) pairs can be reused step
    after step (the sequence numbers live on the device)."""
    tiled, lib = _m("tiled"), _m("_lib")
    rows, own, hl, C, world = 7, 16, 4, 64, 3
    g = torch.Generator().manual_seed(3)
    boxes = [tiled.PeerHalo(2 * rows * hl * C * 4, r, world) for r in range(world)]
    tiled.PeerHalo.connect_local(boxes)
    geo = []                                              # (width, lo, hi) of every rank's strip
    for r in range(world):
        left, right = (hl if r > 0 else 0), (hl if r < world - 1 else 0)
        geo.append((left + own + right, left, left + own))
    slots = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    for step in range(3):
        xs = [torch.randn(1, rows, w, C, generator=g).mul(1 + r + step).cuda() for r, (w, _, _) in enumerate(geo)]
        before = [x.clone() for x in xs]
        L = lib.lib()
        for b in boxes:
            b.begin_step()
        for k in range(2):                                # two exchanges per step, second one of the same tensors again
            for r, b in enumerate(boxes):
                w, lo, hi = geo[r]
                lib.check(L.adpst_halo_push(b._h, b._slot + k, k * rows * hl * C * 4, lib.ptr(xs[r]), rows, w, C, hl, lo, hi,
                                            lib.stream_ptr()))
            for r, b in enumerate(boxes):
                w, lo, hi = geo[r]
                slots[r].zero_()
                lib.check(L.adpst_halo_pull(b._h, b._slot + k, k * rows * hl * C * 4, lib.ptr(xs[r]), rows, w, C, hl, lo, hi,
                                            lib.ptr(slots[r]), lib.stream_ptr()))
        torch.cuda.synchronize()
        for r in range(world):
            w, lo, hi = geo[r]
            assert torch.equal(xs[r][:, :, lo:hi], before[r][:, :, lo:hi])
            got_max = 0.0
            if r > 0:
                _, plo, phi = geo[r - 1]
                assert torch.equal(xs[r][:, :, :lo], before[r - 1][:, :, phi - hl:phi])
                got_max = max(got_max, float(xs[r][:, :, :lo].abs().max()))
            if r < world - 1:
                _, plo, phi = geo[r + 1]
                assert torch.equal(xs[r][:, :, hi:], before[r + 1][:, :, plo:plo + hl])
                got_max = max(got_max, float(xs[r][:, :, hi:].abs().max()))
            assert slots[r].view(torch.float32).item() == got_max
    # argument checks
    with pytest.raises(lib.AdpstError):
        lib.check(lib.lib().adpst_halo_push(boxes[0]._h, 0, 1 << 30, lib.ptr(xs[0]), rows, geo[0][0], C, hl, 0, own, lib.stream_ptr()))
    with pytest.raises(lib.AdpstError):
        lib.check(lib.lib().adpst_halo_push(boxes[0]._h, 99, 0, lib.ptr(xs[0]), rows, geo[0][0], C, hl, 0, own, lib.stream_ptr()))
