"""Shared host logic of the two MattingLaplacian front ends (matting_v2.py / matting_v3.py)."""
import ctypes

import torch

from .. import _lib


def as_cuda_tensor(a, dtype=None):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    if not t.is_cuda:
        _lib.require_cuda()
        t = t.cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class LaplacianHandle:
    """Owns one adpst_laplacian*.  `storage_dtype` is the dtype of image / x / y in HBM, `compute_dtype` the
    arithmetic type of the stencil."""

    KERNELS = {"auto": _lib.LAP_KERNEL_AUTO, "dia": _lib.LAP_KERNEL_DIA, "matrix_free": _lib.LAP_KERNEL_MATRIX_FREE,
               "tile": _lib.LAP_KERNEL_TILE}

    def __init__(self, mode, image, epsilon, window_radius, storage_dtype=None, compute_dtype=None, kernel=None):
        _lib.require_cuda()
        img = as_cuda_tensor(image)
        if img.dim() != 3 or img.shape[2] != 3:
            raise ValueError("image must have shape (H, W, 3), got %s" % (tuple(img.shape),))
        if img.dtype not in (torch.float32, torch.float64):
            raise TypeError("image must be float32 or float64, got %s" % img.dtype)
        self.operator_dtype = img.dtype
        self.storage_dtype = storage_dtype or img.dtype
        self.compute_dtype = compute_dtype or img.dtype
        img = img.to(self.storage_dtype).contiguous()
        self.H, self.W = int(img.shape[0]), int(img.shape[1])
        self.mode, self.radius, self.epsilon = mode, int(window_radius), float(epsilon)
        self.device = img.device
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_laplacian_create(
                mode, self.H, self.W, self.radius, self.epsilon, _lib.ptr(img), _lib.dtype_code(self.storage_dtype),
                _lib.dtype_code(self.compute_dtype), _lib.stream_ptr(), ctypes.byref(h)))
        self._h = h
        self._xlx = torch.zeros(1, dtype=torch.float64, device=self.device)
        if kernel is not None:
            self.set_kernel(kernel)

    def set_kernel(self, kernel):
        """'auto' (default: 'dia' for radius 1 with float32 storage, else 'matrix_free'), 'dia' (precomputed 5x5 stencil
        coefficients, float32 evaluation of L I + L (x - I): the hot path), 'matrix_free' (window statistics recomputed in
        compute_dtype every call: 1e-9 in float64), 'tile' (shared-memory variant of matrix_free; validation)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_laplacian_set_kernel(self._h, self.KERNELS[kernel], _lib.stream_ptr()))

    @property
    def kernel(self):
        k = int(_lib.lib().adpst_laplacian_kernel(self._h))
        return {v: n for n, v in self.KERNELS.items()}[k]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().adpst_laplacian_destroy(h)
            except Exception:
                pass

    def set_quadratic_window(self, col_lo, col_hi):
        """Restrict x^T L x to columns [col_lo, col_hi) (spatially tiled runs); (0, 0) = all columns."""
        _lib.check(_lib.lib().adpst_laplacian_set_quadratic_window(self._h, int(col_lo), int(col_hi)))

    @property
    def HW(self):
        return self.H * self.W

    def apply3(self, x, want_y=True, want_quad=False, y_scale=1.0, out=None, quad_out=None):
        """x: (HW,3) storage_dtype contiguous.  Returns (y or None, xLx 0-dim float64 tensor or None).
        quad_out: optional 1-element float64 device tensor that receives x^T L x instead of the internal one."""
        if x.dtype != self.storage_dtype:
            raise TypeError("x has dtype %s, operator stores %s" % (x.dtype, self.storage_dtype))
        if x.numel() != self.HW * 3:
            raise ValueError("x must have %d x 3 elements, got %s" % (self.HW, tuple(x.shape)))
        y = None
        if want_y:
            y = out if out is not None else torch.empty_like(x)
        q = (quad_out if quad_out is not None else self._xlx) if want_quad else None
        if q is not None and (q.dtype != torch.float64 or q.numel() != 1 or not q.is_cuda):
            raise TypeError("quad_out must be a 1-element float64 CUDA tensor")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_laplacian_matvec(self._h, _lib.ptr(x), _lib.ptr(y), float(y_scale), _lib.ptr(q),
                                                         _lib.stream_ptr()))
        if not want_quad:
            return y, None
        # the internal scalar is overwritten by the next call: hand out a copy unless the caller owns the buffer
        return y, (q.reshape(-1)[0] if quad_out is not None else q.reshape(-1)[0].clone())

    def matmul(self, x):
        """(HW, C') -> (HW, C').  C' != 3 is processed three columns at a time (zero padded)."""
        x = as_cuda_tensor(x)
        if x.dim() != 2 or x.shape[0] != self.HW:
            raise ValueError("x must have shape (%d, C'), got %s" % (self.HW, tuple(x.shape)))
        if x.dtype != self.operator_dtype:
            raise TypeError("x has dtype %s but the operator has dtype %s" % (x.dtype, self.operator_dtype))
        xs = x.to(self.storage_dtype)
        C = x.shape[1]
        if C == 3:
            y, _ = self.apply3(xs.contiguous())
            return y.to(self.operator_dtype)
        outs = []
        for c0 in range(0, C, 3):
            chunk = torch.zeros(self.HW, 3, dtype=self.storage_dtype, device=self.device)
            n = min(3, C - c0)
            chunk[:, :n] = xs[:, c0:c0 + n]
            y, _ = self.apply3(chunk)
            outs.append(y[:, :n])
        return torch.cat(outs, 1).to(self.operator_dtype)
