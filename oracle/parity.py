"""Gradient comparison that knows about ReLU decisions.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's parity leg).

The image gradient of the VGG19 path (style_transfer.py:341 of the reference) is piecewise smooth: a pre-activation
within float32 rounding distance of zero can fall on different sides in a float32 implementation and in the float64
oracle.  Such a flip changes nothing in the forward pass (the activation is ~0 either way) but switches the
back-propagated signal of that ONE unit on or off, which changes the image gradient inside that unit's receptive field
-- and nowhere else.  So the check is: max-norm 1e-5 everywhere outside the receptive fields of the flipped units, a
loose bound inside them, and an explicit count of the flips (any float32 implementation, TF included, has them).
"""
import numpy as np

# (receptive-field size in image pixels, stride in image pixels) of the 13 conv outputs block1_conv1 .. block5_conv1
# (3x3 SAME convolutions, 2x2/2 pools after conv 1, 3, 7, 11):  rf += 2*jump per conv; pool: rf += jump, jump *= 2
RF = []
_rf, _j = 1, 1
for _i in range(13):
    _rf += 2 * _j
    RF.append((_rf, _j))
    if _i in (1, 3, 7, 11):
        _rf += _j
        _j *= 2
del _rf, _j, _i


def flipped_units(acts_gpu, ref_acts):
    """acts_gpu[i] / ref_acts[i]: (1,h,w,C) post-ReLU outputs of conv i (numpy or torch; None entries are skipped).
    Returns a list of (conv index, y, x) of positions where at least one channel's ReLU decision differs, and the total
    number of differing (position, channel) units."""
    pos, total = [], 0
    for i, (a, r) in enumerate(zip(acts_gpu, ref_acts)):
        if a is None or r is None:
            continue
        a = np.asarray(a.detach().cpu() if hasattr(a, "detach") else a)
        r = np.asarray(r.detach().cpu() if hasattr(r, "detach") else r)
        d = (a > 0) != (r > 0)
        total += int(d.sum())
        ys, xs = np.nonzero(d.reshape(d.shape[-3], d.shape[-2], d.shape[-1]).any(-1))
        pos.extend((i, int(y), int(x)) for y, x in zip(ys, xs))
    return pos, total


def influence_mask(flips, H, W):
    """Boolean (H,W): image pixels inside the receptive field of any flipped unit (one stride of slack per side)."""
    m = np.zeros((H, W), dtype=bool)
    for i, y, x in flips:
        rf, j = RF[i]
        half = rf // 2 + j
        cy, cx = y * j + (j - 1) // 2, x * j + (j - 1) // 2
        m[max(0, cy - half):cy + half + 1, max(0, cx - half):cx + half + 1] = True
    return m


def gradient_report(g, g_ref, acts_gpu=None, ref_acts=None, tol=1e-5):
    """g, g_ref: (1,H,W,3) or (H,W,3).  Returns a dict with the max-norm relative error overall, outside / inside the
    receptive fields of flipped ReLU units, the flip count and the fraction of pixels above `tol`."""
    g = np.asarray(g, np.float64).reshape(-1, np.shape(g)[-2], 3)
    go = np.asarray(g_ref, np.float64).reshape(g.shape)
    H, W = g.shape[0], g.shape[1]
    scale = max(float(np.abs(go).max()), 1e-300)
    err = np.abs(g - go).max(-1) / scale
    rep = {"rel_maxnorm": float(err.max()), "frac_pixels_above_tol": float((err > tol).mean()), "relu_flips": 0,
           "rel_maxnorm_outside_flipped_fields": float(err.max()), "pixels_inside_flipped_fields": 0.0}
    if acts_gpu is not None and ref_acts is not None:
        flips, total = flipped_units(acts_gpu, ref_acts)
        mask = influence_mask(flips, H, W)
        rep["relu_flips"] = total
        rep["pixels_inside_flipped_fields"] = float(mask.mean())
        rep["rel_maxnorm_outside_flipped_fields"] = float(err[~mask].max()) if (~mask).any() else 0.0
        rep["rel_maxnorm_inside_flipped_fields"] = float(err[mask].max()) if mask.any() else 0.0
    return rep


def assert_gradient_close(g, g_ref, acts_gpu, ref_acts, tol=1e-5):
    """1e-5 max-norm outside the receptive fields of flipped units; inside them the change is bounded by the flipped unit's
    own back-propagated signal (a few 1e-3 of the largest gradient entry at most)."""
    rep = gradient_report(g, g_ref, acts_gpu, ref_acts, tol)
    if rep["rel_maxnorm"] < tol:
        return rep
    assert rep["relu_flips"] > 0, "gradient off by %.2e with identical ReLU decisions" % rep["rel_maxnorm"]
    assert rep["rel_maxnorm_outside_flipped_fields"] < tol, rep
    assert rep["rel_maxnorm"] < 2e-2, rep
    return rep
