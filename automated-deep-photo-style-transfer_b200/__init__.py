"""B200-native hot path of aRI0U/automated-deep-photo-style-transfer.

Layout
  csrc/            hand-written sm_100a CUDA kernels + the C-ABI (include/adpst.h) -> libadpst.so
  _lib.py          ctypes binding; raises if the library is missing (no CPU fallback)
  components/      host-side mirror of the reference's classes (same module names and signatures):
                   loss.Loss, matting_v2.MattingLaplacian, matting_v3.MattingLaplacian,
                   VGG19.model.StyleContentModel, semantic_merge mask helpers
  style_transfer.py  train_step / optimisation loop (style_transfer.py:295-367 of the reference)
  synth.py         deterministic synthetic inputs for tests and bench

The directory name contains '-', so import it with importlib.import_module("automated-deep-photo-style-transfer_b200")
or through the `adpst_b200` alias module at the repository root.
"""
__version__ = "0.1.0"
