"""Generate tests/golden/*.npz by running the REFERENCE's own code (import only, nothing copied).

Run in the authoring container, where /root/reference exists:   python oracle/make_golden.py
The GPU box has no /root/reference; it only reads the committed .npz files.

What can be executed from the reference without TensorFlow:
  * components/matting_v3.py  compute_laplacian  (numpy/scipy; needs a stub `tensorflow` and np.mat shim)
  * components/semantic_merge.py  extract_segmentation_masks / mask_for_tf / reduce_dict
    (needs stubs for tensorflow, nltk, sematch, components.PSPNet.model, components.util)
Everything else on the path lives in TensorFlow/Keras and stays "parity unpinned".
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


if __name__ == "__main__":
    main()
