"""Drop-in for /root/reference/components/matting_v3.py (class MattingLaplacian, :13-102).

The reference materialises the Levin Laplacian as a COO tf.SparseTensor (81 entries per interior window) and
multiplies with tf.sparse.sparse_dense_matmul.  Here `matmul` is the same operator evaluated matrix-free by the
stencil kernel in "interior windows" mode; the explicit triplets are produced only on demand (`.laplacian`), in
the reference's emission order with duplicates kept, so the sparsity pattern can be compared bit for bit.
"""
import collections

import torch

from .. import _lib
from ._matting_base import LaplacianHandle

SparseCOO = collections.namedtuple("SparseCOO", ["indices", "values", "dense_shape"])   # mirrors tf.SparseTensor


class MattingLaplacian:
    r"""Matting Laplacian of "A closed-form solution to natural image matting" (Levin et al.).
    reference: matting_v3.py:27-39 (constructor), :50-51 (matmul), :61-102 (compute_laplacian)."""

    def __init__(self, image, epsilon=1e-5, window_radius=1, fname=None, *, storage_dtype=None, compute_dtype=None, kernel=None):
        # the reference evaluates compute_laplacian in float64 numpy whatever the image dtype (image.numpy() of a
        # float32 tensor is promoted by np.linalg / einsum only partly); float64 arithmetic is the faithful choice.
        self._op = LaplacianHandle(_lib.LAP_V3, image, epsilon, window_radius, storage_dtype,
                                   compute_dtype or torch.float64, kernel)
        self.size = (self._op.H, self._op.W, 3)                            # :35
        self.dtype = self._op.operator_dtype
        self._coo = None

    @property
    def shape(self):                                                       # :42-48
        H, W, _ = self.size
        return torch.Size((H * W, H * W))

    def matmul(self, x):                                                   # :50-51
        return self._op.matmul(x)

    @property
    def nnz(self):
        return int(_lib.lib().adpst_laplacian_nnz(self._op._h))

    @property
    def laplacian(self):                                                   # :36-39, :97-102
        if self._coo is None:
            op, n = self._op, self.nnz
            rows = torch.empty(n, dtype=torch.int64, device=op.device)
            cols = torch.empty(n, dtype=torch.int64, device=op.device)
            vals = torch.empty(n, dtype=op.storage_dtype, device=op.device)
            with torch.cuda.device(op.device):
                _lib.check(_lib.lib().adpst_laplacian_export_coo(op._h, _lib.ptr(rows), _lib.ptr(cols), _lib.ptr(vals),
                                                                 _lib.stream_ptr()))
            self._coo = SparseCOO(torch.stack([rows, cols], 1), vals.to(self.dtype), self.shape)
        return self._coo

    def quadratic_form(self, x, want_y=False, y_scale=1.0, out=None, quad_out=None):
        x = x.to(self._op.storage_dtype).reshape(-1, 3).contiguous()
        return self._op.apply3(x, want_y=want_y, want_quad=True, y_scale=y_scale, out=out, quad_out=quad_out)
