"""CPU: pin the Laplacian oracle.  v3 against COO triplets emitted by the reference's own
compute_laplacian (tests/golden/v3_*.npz); v2 against v3 on the interior and against an independent
per-window restatement, plus the invariants of SURVEY §4 (symmetric, L.1 = 0, PSD)."""
import numpy as np
import pytest

from conftest import golden
from oracle import matting


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_v3_restatement_matches_reference_coo(tag):
    g = golden("v3_%s.npz" % tag)
    rows, cols, vals, shape = matting.v3_compute_laplacian(g["image"], float(g["eps"]), int(g["r"]))
    assert np.array_equal(rows, g["rows"])          # sparsity pattern and emission order: bit-exact
    assert np.array_equal(cols, g["cols"])
    np.testing.assert_allclose(vals, g["vals"], rtol=0, atol=1e-9 * np.abs(g["vals"]).max())
    H, W, _ = g["image"].shape
    assert len(vals) == 81 * (H - 2) * (W - 2)
    op = matting.V3Operator(g["image"], float(g["eps"]), 1)
    np.testing.assert_allclose(op.matmul(g["x"]), g["Lx"], rtol=1e-9, atol=1e-9 * np.abs(g["Lx"]).max())


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_v2_equals_v3_on_interior(tag):
    g = golden("v3_%s.npz" % tag)
    img = g["image"]
    H, W, _ = img.shape
    op2 = matting.V2Operator(img, float(g["eps"]), 1)
    y2 = op2.matmul(g["x"]).reshape(H, W, 3)
    y3 = g["Lx"].reshape(H, W, 3)
    scale = np.abs(y3).max()
    if H > 4 and W > 4:
        assert np.abs(y2[2:-2, 2:-2] - y3[2:-2, 2:-2]).max() < 1e-9 * scale
    assert np.abs(y2 - y3).max() > 1e-3 * scale      # the border ring differs


@pytest.mark.parametrize("r", [1, 2])
def test_v2_integral_image_equals_direct(r, synth):
    H, W = 9, 11
    img = synth.image(H, W, 5)[0].astype(np.float64)
    x = np.random.default_rng(6).random((H * W, 3))
    op = matting.V2Operator(img, 1e-7, r)
    y = op.matmul(x)
    yd = matting.v2_direct(img, x, 1e-7, r)
    assert np.abs(y - yd).max() < 1e-9 * np.abs(yd).max()


def test_v2_invariants(synth):
    H, W = 7, 8
    img = synth.image(H, W, 7)[0].astype(np.float64)
    op = matting.V2Operator(img, 1e-5, 1)
    L = op.matmul(np.eye(H * W))
    assert np.abs(L - L.T).max() < 1e-10
    assert np.abs(L @ np.ones(H * W)).max() < 1e-9
    assert np.linalg.eigvalsh((L + L.T) / 2).min() > -1e-9
    assert op.means.shape == (H, W, 3, 1) and op.delta_inv.shape == (H, W, 3, 3)
    assert op.shape == (H * W, H * W)


def test_v2_x_equals_image_is_tiny(synth):
    """x = I (iteration 0): Lx is ~eps-small; this is the case that needs float64."""
    img = synth.image(16, 16, 8)[0].astype(np.float64)
    op = matting.V2Operator(img, 1e-7, 1)
    y = op.matmul(img.reshape(-1, 3))
    assert np.abs(y).max() < 1e-4


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_v2_restatement_matches_reference_code(tag):
    """tests/golden/v2_*.npz come from the reference's OWN matting_v2.py (its __init__ and _matmul), executed over the numpy
    stand-in for its TensorFlow primitives (oracle/tf_shim.py, oracle/make_golden.py): fields, shape and mat-vecs."""
    g = golden("v2_%s.npz" % tag)
    op = matting.V2Operator(g["image"], float(g["eps"]), int(g["r"]))
    assert tuple(op.shape) == tuple(g["shape"])
    np.testing.assert_allclose(op.means, g["means"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(op.delta_inv, g["delta_inv"], rtol=1e-10, atol=0)
    for x, y in ((g["x"], g["Lx"]), (g["image"].reshape(-1, 3), g["LI"])):
        np.testing.assert_allclose(op.matmul(x), y, rtol=0, atol=1e-9 * max(np.abs(g["Lx"]).max(), 1e-30))


def test_operator_is_a_symmetric_5x5_stencil_with_zero_row_sums(synth):
    """The facts the diagonal-format kernel rests on, checked on the ORACLE (dense matrix of the float64 restatement):
    L = L^T, footprint 5x5, L 1 = 0 -- for v2 (symmetric padding folds back inside) and v3."""
    H, W = 9, 11
    img = synth.smooth_image(16, 16, 3)[0][:H, :W].astype(np.float64)
    idx = np.arange(H * W)
    yy, xx = idx // W, idx % W
    far = (np.abs(yy[:, None] - yy[None, :]) > 2) | (np.abs(xx[:, None] - xx[None, :]) > 2)
    for op in (matting.V2Operator(img, 1e-7, 1), matting.V3Operator(img, 1e-7, 1)):
        A = np.asarray(op.matmul(np.eye(H * W)))
        s = np.abs(A).max()
        assert np.abs(A - A.T).max() < 1e-9 * s and np.abs(A[far]).max() < 1e-9 * s and np.abs(A.sum(1)).max() < 1e-9 * s


def test_golden_provenance_is_recorded():
    """tests/golden/PROVENANCE.json names the reference files whose own code produced the vectors (sha256); where the
    reference checkout is present (the authoring container) the hashes must still match."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    rec = json.load(open(os.path.join(GOLDEN, "PROVENANCE.json")))
    assert set(rec["executed_files"]) == {"components/matting_v3.py", "components/semantic_merge.py",
                                          "components/matting_v2.py", "components/loss.py"}
    assert set(rec["golden_files"]) == {f for f in os.listdir(GOLDEN) if f.endswith(".npz")}
    if os.path.isdir(rec["reference_root"]):
        for rel, digest in rec["executed_files"].items():
            with open(os.path.join(rec["reference_root"], rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, rel


def test_extended_precision_restatement_agrees_with_the_oracle(synth=None):
    """oracle.matting.v2_extended_precision (np.longdouble, centred moments, Cholesky) against the float64 restatements:
    identical to ~1e-14 for generic x; at x = I on a grey (rank-1 covariance) image the float64 oracle itself is only good
    to ~1e-5 relative on I^T L I, which is why the GPU tests judge that case against the extended-precision version."""
    import importlib
    from conftest import PKG_NAME
    synth = importlib.import_module(PKG_NAME + ".synth")
    img = synth.image(7, 9, 3)[0]
    x = np.random.default_rng(0).random((63, 3))
    a = np.asarray(matting.v2_extended_precision(img, x, 1e-7), np.float64)
    b = matting.V2Operator(img.astype(np.float64), 1e-7, 1).matmul(x)
    assert np.abs(a - b).max() < 1e-11 * np.abs(b).max()
    g = np.ascontiguousarray(np.repeat(synth.smooth_image(8, 8, 44)[0][:2, :3, :1], 3, -1))
    I = g.reshape(-1, 3)
    t = matting.v2_extended_precision(g, I, 1e-7)
    o = matting.V2Operator(g.astype(np.float64), 1e-7, 1).matmul(I.astype(np.float64))
    qt, qo = float((np.asarray(I, np.longdouble) * t).sum()), float((I.astype(np.float64) * o).sum())
    assert qt > 0 and abs(qo - qt) < 1e-3 * qt          # same quantity ...
    print("I^T L I: extended %.12e, float64 oracle %.12e (relative difference %.1e)" % (qt, qo, abs(qo - qt) / qt))
