"""CPU oracle for the per-iteration loss-and-gradient path of style_transfer.py.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`automated-deep-photo-style-transfer_b200/`) imports this.  Allowed importers:
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Parity status (see DESIGN.md "Oracle and parity status"); golden vectors are made by oracle/make_golden.py, which runs the
reference's own code in the authoring container:
  * matting_v3 build  -- PINNED: COO triplets produced by the reference's own `compute_laplacian`
    (tests/golden/v3_*.npz).
  * mask extraction / class order -- PINNED the same way (tests/golden/masks_*.npz).
  * matting_v2 fields + matvec (r = 1, 2, 3) -- PINNED to the reference's code: its own matting_v2.py executed over
    oracle/tf_shim.py, a numpy stand-in for the TensorFlow primitives it calls (tests/golden/v2_*.npz).  Rests on
    tf.pad(SYMMETRIC) == np.pad('symmetric').
  * Loss (content, masked Gram, style, photorealism, total, key order) -- PINNED to the reference's code: its own loss.py
    over the same shim (tests/golden/loss_*.npz).  Rests on the published tf.image.resize bilinear formula.
  * VGG19 arithmetic -- Keras is un-vendored and not installable: checked against torchvision's VGG19 (an independent
    implementation of the same network); "parity unpinned" against Keras itself.
  * Adam + clip -- restatement of documented Keras behaviour: "parity unpinned".
"""
