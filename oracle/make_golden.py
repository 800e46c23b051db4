"""Generate tests/golden/*.npz by running the REFERENCE's own code (import only, nothing copied).

Run in the authoring container, where /root/reference exists:   python oracle/make_golden.py
The GPU box has no /root/reference; it only reads the committed .npz files.

What can be executed from the reference without TensorFlow:
  * components/matting_v3.py  compute_laplacian  (numpy/scipy; needs a stub `tensorflow` and np.mat shim)
  * components/semantic_merge.py  extract_segmentation_masks / mask_for_tf / reduce_dict
    (needs stubs for tensorflow, nltk, sematch, components.PSPNet.model, components.util)
  * components/matting_v2.py and components/loss.py: their own code over oracle/tf_shim.py, a numpy stand-in for the
    few TensorFlow primitives they call (the shim's header says which two of those rest on documented TF behaviour)
The VGG19 arithmetic (Keras) and the optimiser (tf.optimizers.Adam) cannot be executed and stay "parity unpinned".

SECURITY NOTE: this script exec()s third-party Python from /root/reference (untrusted public content) inside the calling
process, with the caller's privileges.  Run it ONLY in a throw-away sandbox without network access or credentials (the
authoring container is one).  Nothing under tests/ or the product package ever imports the reference; the committed .npz
files are plain arrays.  `python oracle/make_golden.py --provenance` (re)writes tests/golden/PROVENANCE.json with the sha256
of every reference file that was executed, so the origin of the vectors can be audited against a given checkout.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _stub_modules():
    tf = types.ModuleType("tensorflow")
    tf.linalg = types.SimpleNamespace(LinearOperator=object)
    tf.SparseTensor = lambda indices, data, shape: (np.asarray(indices), np.asarray(data), tuple(shape))
    tf.constant = lambda a, *k, **kw: np.asarray(a)
    tf.expand_dims = lambda a, axis: np.expand_dims(a, axis)
    sys.modules["tensorflow"] = tf
    nltk = types.ModuleType("nltk")
    nltk.data = types.SimpleNamespace(path=[])
    sys.modules["nltk"] = nltk
    for name in ("sematch", "sematch.semantic", "sematch.semantic.similarity"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["sematch.semantic.similarity"].WordNetSimilarity = lambda: None
    for name in ("components", "components.PSPNet", "components.PSPNet.model", "components.util"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["components.PSPNet.model"].load_color_label_dict = lambda: {}
    sys.modules["components.util"].WEIGHTS_DIR = "/nonexistent"
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix                      # matting_v3.py:101 uses np.mat (removed in NumPy 2)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


EXECUTED_REFERENCE_FILES = ["components/matting_v3.py", "components/semantic_merge.py", "components/matting_v2.py",
                            "components/loss.py"]


def write_provenance():
    """sha256 of the reference files this script executes -> tests/golden/PROVENANCE.json."""
    import hashlib
    import json
    rec = {"reference_root": REF, "generator": "oracle/make_golden.py", "executed_files": {}}
    for rel in EXECUTED_REFERENCE_FILES:
        with open(os.path.join(REF, rel), "rb") as f:
            rec["executed_files"][rel] = hashlib.sha256(f.read()).hexdigest()
    rec["golden_files"] = sorted(f for f in os.listdir(OUT) if f.endswith(".npz"))
    with open(os.path.join(OUT, "PROVENANCE.json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    return rec


def main():
    _stub_modules()
    os.makedirs(OUT, exist_ok=True)
    v3 = _load(os.path.join(REF, "components", "matting_v3.py"), "ref_matting_v3")
    sm = _load(os.path.join(REF, "components", "semantic_merge.py"), "ref_semantic_merge")
    inst = object.__new__(v3.MattingLaplacian)

    # --- v3 COO triplets from the reference's compute_laplacian --------------------------------
    cases = [("a", 6, 7, 1e-7, 11, "uniform"), ("b", 9, 12, 1e-5, 12, "uniform"),
             ("c", 16, 13, 1e-7, 13, "smooth")]
    from importlib import import_module
    synth = import_module("automated-deep-photo-style-transfer_b200.synth")
    for tag, H, W, eps, seed, kind in cases:
        img = (synth.image(H, W, seed) if kind == "uniform" else synth.smooth_image(H, W, seed, passes=2))[0]
        img = img.astype(np.float64)
        idx, data, shape = inst.compute_laplacian(img, eps=eps, win_rad=1)
        idx = np.asarray(idx)
        x = np.random.default_rng(seed + 100).random((H * W, 3))
        import scipy.sparse
        L = scipy.sparse.coo_matrix((data, (idx[:, 0], idx[:, 1])), shape=shape).tocsr()
        np.savez_compressed(os.path.join(OUT, "v3_%s.npz" % tag), image=img, eps=eps, r=1,
                            rows=idx[:, 0].astype(np.int64), cols=idx[:, 1].astype(np.int64),
                            vals=np.asarray(data, np.float64), x=x, Lx=L @ x, LI=L @ img.reshape(H * W, 3))
        print("v3_%s: %dx%d nnz=%d" % (tag, H, W, len(data)))

    # --- mask helpers ---------------------------------------------------------------------------
    for tag, H, W, K, seed, cell in [("a", 24, 40, 4, 3, 8), ("b", 33, 21, 7, 4, 5)]:
        seg = synth.label_image(H, W, K, seed, cell=cell)
        d = sm.extract_segmentation_masks(seg)
        keys = sorted(d)
        masks = sm.mask_for_tf(d)
        back = sm.reduce_dict(d, np.zeros((1, H, W, 3)))
        np.savez_compressed(os.path.join(OUT, "masks_%s.npz" % tag), seg=seg,
                            keys=np.array(keys, dtype=np.int64),
                            masks=np.stack([np.asarray(m) for m in masks]).astype(np.float32),
                            reduced=back.astype(np.int64))
        print("masks_%s: K=%d keys=%s" % (tag, len(keys), keys))


def main_tf_shim():
    """Golden vectors from the reference's own matting_v2.py and loss.py, executed over oracle/tf_shim.py."""
    import argparse
    from importlib import import_module
    from oracle import tf_shim
    synth = import_module("automated-deep-photo-style-transfer_b200.synth")
    sys.modules["tensorflow"] = tf_shim.make_module()
    for name in ("components", "components.NIMA", "components.NIMA.model"):
        sys.modules[name] = types.ModuleType(name)
    # NIMA is out of scope (weight 0): a stand-in that returns uniform scores keeps compute_loss runnable
    sys.modules["components.NIMA.model"].NIMAModel = lambda: (lambda image: np.full((1, 10), 0.1, dtype=image.dtype))
    v2 = _load(os.path.join(REF, "components", "matting_v2.py"), "components.matting_v2")
    sys.modules["components.matting_v2"] = v2
    loss_mod = _load(os.path.join(REF, "components", "loss.py"), "ref_loss")

    # --- v2 operator: fields and mat-vecs ---------------------------------------------------------
    for tag, H, W, eps, r, seed, kind in [("a", 7, 9, 1e-7, 1, 21, "uniform"), ("b", 12, 10, 1e-5, 1, 22, "smooth"),
                                          ("c", 11, 13, 1e-5, 3, 23, "uniform"), ("d", 9, 8, 1e-7, 2, 24, "smooth")]:
        img32 = (synth.image(H, W, seed) if kind == "uniform" else synth.smooth_image(H, W, seed, passes=2))[0]
        img = img32.astype(np.float64)                         # float32-representable values, float64 arithmetic
        op = v2.MattingLaplacian(img, epsilon=eps, window_radius=r)
        x = np.random.default_rng(seed + 100).random((H * W, 3)).astype(np.float32).astype(np.float64)
        np.savez_compressed(os.path.join(OUT, "v2_%s.npz" % tag), image=img, eps=eps, r=r, x=x,
                            Lx=np.asarray(op.matmul(x)), LI=np.asarray(op.matmul(img.reshape(H * W, 3))),
                            means=np.asarray(op.means), delta_inv=np.asarray(op.delta_inv),
                            shape=np.asarray(op.shape, dtype=np.int64))
        print("v2_%s: %dx%d r=%d" % (tag, H, W, r))

    # --- Loss: content, masked-Gram style, photorealism, weighted total ------------------------------
    # (inputs are float32-representable but handed over as float64, so the reference's arithmetic runs in float64 and the
    #  vectors pin the logic to ~1e-15 instead of float32 round-off)
    for tag, H, W, K, seed, photo_w in [("a", 16, 16, 3, 31, 1e4), ("b", 16, 24, 0, 32, 0.0)]:
        rng = np.random.default_rng(seed)
        image = synth.image(H, W, seed).astype(np.float64)                  # (1,H,W,3)
        shapes = {"block1": (H, W, 64), "block2": (H // 2, W // 2, 64), "block3": (H // 4, W // 4, 128)}
        sshapes = {"block1": (H + 4, W - 6, 64), "block2": (H // 2 + 2, W // 2 - 3, 64), "block3": (H // 4 + 1, W // 4 - 1, 128)}
        feat = lambda shp: (rng.random((1,) + shp) * 20.0).astype(np.float32).astype(np.float64)
        content_layers, style_layers = ["block3"], ["block1", "block2", "block3"]
        content_target = {n: feat(shapes[n]) for n in content_layers}
        style_target = {n: feat(sshapes[n]) for n in style_layers}
        outputs = {"content": {n: feat(shapes[n]) for n in content_layers}, "style": {n: feat(shapes[n]) for n in style_layers}}
        cm = sm = None
        if K:
            cmask = synth.label_image(H, W, K, seed + 1, cell=4)
            smask = synth.label_image(H + 4, W - 6, K, seed + 2, cell=4)
            msk = import_module("oracle.masks")
            cm = [np.asarray(m, np.float64) for m in msk.mask_for_tf(msk.extract_segmentation_masks(cmask))]
            sm = [np.asarray(m, np.float64) for m in msk.mask_for_tf(msk.extract_segmentation_masks(smask))]
            assert len(cm) == len(sm) == K
        args = argparse.Namespace(content_weight=1.0, style_weight=100.0, nima_weight=0.0, regularization_weight=photo_w,
                                  matting_epsilon=1e-7, matting_window_radius=1)
        L = loss_mod.Loss(content_target, style_target, args, cm, sm)
        if photo_w > 0:
            L.initialize_matting_laplacian(image[0])
        d = L(image, outputs)
        out = {"image": image.astype(np.float32), "K": K, "photo_weight": photo_w,
               "content_loss": np.float64(d["Content loss"]), "style_loss": np.float64(d["Style loss"]),
               "total_minus_nima": np.float64(d["Total loss"]) - 0.0 * np.float64(d["NIMA loss"]),
               "keys": np.array(list(d.keys()))}
        if photo_w > 0:
            out["photo_loss"] = np.float64(d["Photorealism regualarization"])
        for n in content_layers:
            out["ct_" + n] = content_target[n].astype(np.float32); out["co_" + n] = outputs["content"][n].astype(np.float32)
        for n in style_layers:
            out["st_" + n] = style_target[n].astype(np.float32); out["so_" + n] = outputs["style"][n].astype(np.float32)
        if K:
            out["cmasks"] = np.stack(cm).astype(np.float32); out["smasks"] = np.stack(sm).astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "loss_%s.npz" % tag), **out)
        print("loss_%s:" % tag, {k: float(v) for k, v in d.items()})


if __name__ == "__main__":
    if "--provenance" in sys.argv:
        print(write_provenance())
    else:
        main()
        main_tf_shim()
        write_provenance()
