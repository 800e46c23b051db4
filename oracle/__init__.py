"""CPU oracle for the per-iteration loss-and-gradient path of style_transfer.py.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`automated-deep-photo-style-transfer_b200/`) imports this.  Allowed importers:
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Parity status (see DESIGN.md "Oracle"):
  * matting_v3 build  -- PINNED: checked against COO triplets produced by the reference's own
    `compute_laplacian` (tests/golden/v3_*.npz, made by oracle/make_golden.py).
  * mask extraction / class order -- PINNED the same way (tests/golden/masks_*.npz).
  * matting_v2 build + matvec -- restatement of matting_v2.py:24-52,147-251; TensorFlow is not
    installable here, so it is pinned only indirectly (equals the pinned v3 operator on the interior
    to 1e-12, symmetric, L.1 = 0).  "parity unpinned" against TF itself.
  * VGG19 / Gram / content / Adam -- restatement of loss.py, VGG19/model.py, style_transfer.py:321-343
    on torch-CPU float64.  The arithmetic lives in TensorFlow/Keras (un-vendored, unpinned version):
    "parity unpinned".
"""
