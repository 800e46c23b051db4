"""Drop-in for /root/reference/components/matting_v2.py (class MattingLaplacian, :6-251).

Same constructor, same attributes (`radius`, `size`, `window_area`, `means`, `delta_inv`, `shape`) and
`matmul(x)`; tensors are torch CUDA tensors instead of tf.Tensor.  The operator is matrix-free on the GPU
(csrc/laplacian.cu); nothing here computes on the host.
"""
import torch

from .. import _lib
from ._matting_base import LaplacianHandle, as_cuda_tensor


class MattingLaplacian:
    r"""Matting Laplacian of "Fast matting using large kernel matting Laplacian matrices" (He et al.),
    symmetric padding, one window per pixel.  reference: matting_v2.py:11-52 (build), :147-176 (matmul)."""

    def __init__(self, image, epsilon=1e-5, window_radius=1, *, storage_dtype=None, compute_dtype=None, kernel=None):
        """image: (H,W,3) float64 (script, style_transfer.py:315) or float32 (benchmark.py:25).
        The operator dtype is image.dtype, as in the reference (:24-25).  `storage_dtype` / `compute_dtype`
        are extensions: float32 HBM traffic with float64 arithmetic is what Loss uses on the hot path."""
        self._op = LaplacianHandle(_lib.LAP_V2, image, epsilon, window_radius, storage_dtype, compute_dtype, kernel)
        self.radius = int(window_radius)                                   # :33
        self.size = (self._op.H, self._op.W, 3)                            # :35
        self.window_area = (2 * self.radius + 1) ** 2                      # :44
        self.dtype = self._op.operator_dtype
        self._coeffs = None

    # LinearOperator surface used by loss.py:159-161 and benchmark.py
    @property
    def shape(self):                                                       # :56-62
        H, W, _ = self.size
        return torch.Size((H * W, H * W))

    def matmul(self, x):                                                   # :147-176
        return self._op.matmul(x)

    def _ensure_coeffs(self):
        if self._coeffs is None:
            op = self._op
            means = torch.empty(op.H, op.W, 3, 1, dtype=op.storage_dtype, device=op.device)
            dinv = torch.empty(op.H, op.W, 3, 3, dtype=op.storage_dtype, device=op.device)
            with torch.cuda.device(op.device):
                _lib.check(_lib.lib().adpst_laplacian_coefficients(op._h, _lib.ptr(means), _lib.ptr(dinv),
                                                                   _lib.stream_ptr()))
            self._coeffs = (means.to(self.dtype), dinv.to(self.dtype))
        return self._coeffs

    @property
    def means(self):                                                       # :51  (H,W,3,1)
        return self._ensure_coeffs()[0]

    @property
    def delta_inv(self):                                                   # :52  (H,W,3,3)
        return self._ensure_coeffs()[1]

    # extension used by Loss: x^T L x (float64) and optionally y_scale * L x in one pass over HBM
    def quadratic_form(self, x, want_y=False, y_scale=1.0, out=None, quad_out=None):
        x = as_cuda_tensor(x, self._op.storage_dtype).reshape(-1, 3)
        return self._op.apply3(x, want_y=want_y, want_quad=True, y_scale=y_scale, out=out, quad_out=quad_out)
