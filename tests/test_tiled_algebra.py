"""CPU: the halo algebra of the spatially tiled run (tiled.py), checked in float64 with torch autograd and no GPU.

A VGG19-shaped stack (same 13 convolutions / ReLUs / four 2x2 max-pools, few channels, random weights) is evaluated (a) on the
whole image and (b) on `world` column strips with the geometry of tiled.Tile -- per-level halos LEVEL_HALO, the pooled tensor
of a strip written into the wider next-level tensor at Tile.pool_xoff, halo columns OVERWRITTEN with the neighbour's own columns
between the SEGMENTS on the way up and, for the gradients, on the way down.  Claim of tiled.py's header: with halo >= 2 x
(convolutions per segment) every own column of every tapped activation and of d(loss)/d(image) is exactly what the whole-image
evaluation gives.  The last test shows that the rule is tight: one column less on block3's level breaks it.
"""
import importlib

import pytest
import torch
import torch.nn.functional as F

from conftest import PKG_NAME

tiled = importlib.import_module(PKG_NAME + ".tiled")
vgg = importlib.import_module(PKG_NAME + ".components.VGG19.model")

CH = (3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 5)          # channels: image, then the 13 convolutions
TAPS = (0, 2, 4, 8, 9, 12)                               # the style / content layers of the real network


def _weights(seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(CH[i + 1], CH[i], 3, 3, generator=g, dtype=torch.float64) * 0.4,
             torch.randn(CH[i + 1], generator=g, dtype=torch.float64) * 0.1) for i in range(13)]


def _level(i):
    return sum(1 for p in vgg.POOL_AFTER if p < i)


def _whole(img, ws, seeds):
    """acts of all convolutions, and d(sum_i <seed_i, act_i>)/d(img); img (1, C, H, W)."""
    x = img.clone().requires_grad_(True)
    acts, cur = [], x
    for i, (w, b) in enumerate(ws):
        cur = F.relu(F.conv2d(cur, w, b, padding=1))
        acts.append(cur)
        if i in vgg.POOL_AFTER:
            cur = F.max_pool2d(cur, 2)
    tot = sum((acts[i] * seeds[i]).sum() for i in TAPS)
    (g,) = torch.autograd.grad(tot, x)
    return [a.detach() for a in acts], g


def _exchange(tensors, tiles):
    """Overwrite every strip's halo columns with the neighbours' own columns (tensors: one (1, C, h, w_l) per rank)."""
    new = [t.clone() for t in tensors]
    for r, t in enumerate(tiles):
        w = tensors[r].shape[3]
        lo, hi = t.own_cols(w)
        hl = t.halo_cols(w)
        if t.has_left:
            plo, phi = tiles[r - 1].own_cols(tensors[r - 1].shape[3])
            new[r][..., lo - hl:lo] = tensors[r - 1][..., phi - hl:phi]
        if t.has_right:
            plo, phi = tiles[r + 1].own_cols(tensors[r + 1].shape[3])
            new[r][..., hi:hi + hl] = tensors[r + 1][..., plo:plo + hl]
    return new


def _strips(img, ws, seeds, world, halos):
    """The same quantities from `world` strips.  Returns (per rank: acts, tile), stitched d/d(image)."""
    W = img.shape[3]
    tiles = [tiled.Tile(W, r, world, halos=halos) for r in range(world)]

    def crop(full, t, level):                    # this rank's columns of a whole-image tensor of that level
        a = (t.own_lo >> level) - t.left[level]
        return full[..., a:a + t.widths[level]]

    seeds_loc = [{i: crop(seeds[i], t, _level(i)) for i in TAPS} for t in tiles]
    # ---- up: segment by segment, every segment a differentiable function of its (detached, exchanged) input
    inputs = [[crop(img, t, 0).clone().requires_grad_(True) for t in tiles]]          # per segment: its input on every rank
    outputs, acts = [], [[None] * 13 for _ in tiles]
    for s, (first, last) in enumerate(vgg.SEGMENTS):
        outs = []
        for r, t in enumerate(tiles):
            cur = inputs[s][r]
            for i in range(first, last + 1):
                cur = F.relu(F.conv2d(cur, ws[i][0], ws[i][1], padding=1))
                acts[r][i] = cur
            if last in vgg.POOL_AFTER:            # the pool of a strip lands inside the wider next-level tensor
                l = _level(last)
                pooled = F.max_pool2d(cur, 2)
                wide = torch.zeros(*pooled.shape[:3], t.widths[l + 1], dtype=pooled.dtype)
                cur = torch.cat([wide[..., :t.pool_xoff[l]], pooled,
                                 wide[..., t.pool_xoff[l] + pooled.shape[3]:]], dim=3)
            outs.append(cur)
        outputs.append(outs)
        if last < 12:
            inputs.append([o.detach().clone().requires_grad_(True) for o in _exchange([o.detach() for o in outs], tiles)])
    # ---- down: the gradient w.r.t. a segment's output is exact on the own columns only; exchange, then continue
    grad_out = [None] * world
    for s in reversed(range(len(vgg.SEGMENTS))):
        first, last = vgg.SEGMENTS[s]
        grads = []
        for r, t in enumerate(tiles):
            tot = sum((acts[r][i] * seeds_loc[r][i]).sum() for i in TAPS if first <= i <= last)
            if grad_out[r] is not None:
                tot = tot + (outputs[s][r] * grad_out[r]).sum()
            (g,) = torch.autograd.grad(tot, inputs[s][r])
            grads.append(g)
        grad_out = _exchange(grads, tiles) if s > 0 else grads
    own = [g[..., t.own_cols(g.shape[3])[0]:t.own_cols(g.shape[3])[1]] for g, t in zip(grad_out, tiles)]
    return [([a.detach() for a in acts[r]], tiles[r]) for r in range(world)], torch.cat(own, dim=3)


def _setup(world, H=32, seed=0):
    W = 64 * world
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, H, W, generator=g, dtype=torch.float64)
    ws = _weights(seed + 1)
    acts, _ = _whole(img, ws, {i: 0.0 for i in TAPS})
    seeds = {i: torch.randn(acts[i].shape, generator=g, dtype=torch.float64) for i in TAPS}
    return img, ws, seeds


@pytest.mark.parametrize("world", [2, 3])
def test_strips_reproduce_the_whole_image(world):
    img, ws, seeds = _setup(world)
    acts, grad = _whole(img, ws, seeds)
    per_rank, grad_strips = _strips(img, ws, seeds, world, tiled.LEVEL_HALO)
    for a_loc, t in per_rank:
        for i in range(13):
            l = _level(i)
            lo, hi = t.own_cols(a_loc[i].shape[3])
            assert a_loc[i].shape[3] == t.widths[l]
            assert torch.equal(a_loc[i][..., lo:hi], acts[i][..., t.own_lo >> l:t.own_hi >> l]), (t.rank, i)
    assert float(grad.abs().max()) > 0
    assert float((grad_strips - grad).abs().max()) <= 1e-12 * float(grad.abs().max())


def test_ragged_height_and_four_strips():
    """H not a multiple of 16 (VALID pooling drops the odd rows), four strips: two interior ranks with halos on both sides."""
    img, ws, seeds = _setup(4, H=44, seed=3)
    acts, grad = _whole(img, ws, seeds)
    per_rank, grad_strips = _strips(img, ws, seeds, 4, tiled.LEVEL_HALO)
    for a_loc, t in per_rank:
        lo, hi = t.own_cols(a_loc[12].shape[3])
        assert torch.equal(a_loc[12][..., lo:hi], acts[12][..., t.own_lo >> 4:t.own_hi >> 4])
    assert float((grad_strips - grad).abs().max()) <= 1e-12 * float(grad.abs().max())


def test_every_rank_plans_the_same_exchanges():
    """The mailbox offsets of tiled.PeerHalo are running sums over exchange_plan: it must not depend on the rank."""
    plans = [tiled.exchange_plan(2160, tiled.Tile(3840, r, 8), 12) for r in range(8)]
    assert all(p == plans[0] for p in plans) and len(plans[0]) == 11


def test_uniform_halos_are_a_special_case():
    """halo 32 / 2^l on level l (what a geometry without per-level offsets needs for block5's two columns)."""
    img, ws, seeds = _setup(2)
    _, grad = _whole(img, ws, seeds)
    _, grad_strips = _strips(img, ws, seeds, 2, (32, 16, 8, 4, 2))
    assert float((grad_strips - grad).abs().max()) <= 1e-12 * float(grad.abs().max())


def test_the_halo_rule_is_tight(monkeypatch):
    """Block3 has four convolutions between two exchanges: 8 halo columns are needed, 6 are not enough (Tile refuses them; with
    the check disabled the gradient near the strip boundary is wrong)."""
    img, ws, seeds = _setup(2)
    _, grad = _whole(img, ws, seeds)
    with pytest.raises(ValueError):
        tiled.Tile(128, 0, 2, halos=(4, 4, 6, 4, 2))
    monkeypatch.setattr(tiled, "_level_convs", lambda: [[0], [0], [0], [0, 0], [0]])
    _, grad_strips = _strips(img, ws, seeds, 2, (4, 4, 6, 4, 2))
    assert float((grad_strips - grad).abs().max()) > 1e-6 * float(grad.abs().max())
