"""Mirror of /root/reference/benchmark.py (:11-45): preprocessing time of the two MattingLaplacian variants.

The reference times `ML2(img, epsilon=1e-5, window_radius=1)` (linear operator: window means + inverse covariances,
matting_v2.py:11-52) against `ML3(...)` (explicit COO matrix, matting_v3.py:27-39,61-102) on float32 U[0,1) images of
50x50 ... 500x500 with timeit.repeat(repeat=5, number=1) and plots mean +- 2 std/sqrt(5) with matplotlib.

Here the same sweep runs on the GPU operators.  Both are matrix-free, so their constructors only copy the image; to time the
work the reference's constructors do, each sample also materialises what the reference keeps: `.means` / `.delta_inv`
for v2, the `.laplacian` COO triplets for v3.  matplotlib is not required: the result is printed as a table and as one
JSON line (and returned by `run`).

    python -m automated-deep-photo-style-transfer_b200.benchmark            (or: python benchmark.py from the package directory)
"""
import json
import timeit

import numpy as np
import torch

try:
    from .components.matting_v2 import MattingLaplacian as ML2
    from .components.matting_v3 import MattingLaplacian as ML3
except ImportError:                                        # executed as a script from the package directory
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    _pkg = os.path.basename(os.path.dirname(os.path.abspath(__file__)))
    ML2 = importlib.import_module(_pkg + ".components.matting_v2").MattingLaplacian
    ML3 = importlib.import_module(_pkg + ".components.matting_v3").MattingLaplacian

eps, r = 1e-5, 1                                           # benchmark.py:11
n_iters = 10                                               # benchmark.py:12
repeats = 5                                                # benchmark.py:14


def _build_v2(img):
    op = ML2(img, epsilon=eps, window_radius=r)
    op.means, op.delta_inv                                 # what matting_v2.py:49-52 computes in the constructor
    torch.cuda.synchronize()
    return op


def _build_v3(img):
    op = ML3(img, epsilon=eps, window_radius=r)
    op.laplacian                                           # what matting_v3.py:36-39 computes in the constructor
    torch.cuda.synchronize()
    return op


def run(sizes=None, repeat=repeats, seed=0, verbose=True):
    """sizes: list of (H, W); default 50x50 ... 500x500 as in the reference.  Returns a dict with per-size mean/std seconds."""
    if not torch.cuda.is_available():
        raise RuntimeError("benchmark needs a CUDA device; there is no CPU fallback")
    if sizes is None:
        sizes = [(50 * (i + 1), 50 * (i + 1)) for i in range(n_iters)]
    g = torch.Generator(device="cuda").manual_seed(seed)
    ticks, t2, t3 = [], [], []
    _build_v2(torch.rand(16, 16, 3, device="cuda", generator=g)); _build_v3(torch.rand(16, 16, 3, device="cuda", generator=g))
    for i, (H, W) in enumerate(sizes):
        tick = '{}x{}'.format(H, W)
        ticks.append(tick)
        msg = '[{}/{}] Evaluating preprocessing time with image of size {}...'.format(i + 1, len(sizes), tick)
        img = torch.rand(H, W, 3, device="cuda", generator=g)                   # tf.random.uniform((H,W,3)), benchmark.py:25
        if verbose:
            print(msg + ' (1/2)', end='\r')
        t2.append(timeit.repeat(lambda: _build_v2(img), repeat=repeat, number=1))
        if verbose:
            print(msg + ' (2/2)', end='\r')
        t3.append(timeit.repeat(lambda: _build_v3(img), repeat=repeat, number=1))
        if verbose:
            print(msg + ' Done.')
    t2, t3 = np.asarray(t2), np.asarray(t3)
    out = {"sizes": ticks, "repeats": repeat, "epsilon": eps, "window_radius": r,
           "linear_operator_v2": {"mean_s": t2.mean(1).tolist(), "std_s": t2.std(1).tolist(),
                                  "bound_s": (2 * t2.std(1) / np.sqrt(repeat)).tolist()},
           "matrix_v3": {"mean_s": t3.mean(1).tolist(), "std_s": t3.std(1).tolist(),
                         "bound_s": (2 * t3.std(1) / np.sqrt(repeat)).tolist()}}
    if verbose:
        print("%-10s %18s %18s" % ("Image size", "Linear operator (s)", "Matrix (s)"))
        for k, tick in enumerate(ticks):
            print("%-10s %10.6f +-%.6f %10.6f +-%.6f" % (tick, out["linear_operator_v2"]["mean_s"][k],
                                                         out["linear_operator_v2"]["bound_s"][k],
                                                         out["matrix_v3"]["mean_s"][k], out["matrix_v3"]["bound_s"][k]))
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
