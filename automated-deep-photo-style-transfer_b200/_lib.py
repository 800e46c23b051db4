"""ctypes binding of include/adpst.h.  Loads libadpst.so from this directory and fails loudly if it is missing:
there is no CPU or eager-PyTorch fallback for any compute entry point."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadpst.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
F32, F64 = 0, 1
LAP_V2, LAP_V3 = 2, 3
LAP_KERNEL_AUTO, LAP_KERNEL_DIA, LAP_KERNEL_MATRIX_FREE, LAP_KERNEL_TILE = 0, 1, 2, 3
VGG_NUM_CONV, VGG_NUM_POOL = 13, 4


class AdpstError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libadpst error %d: %s" % (code, msg))
        self.code = code


_c = ctypes
_vp, _i, _d, _f, _sz = _c.c_void_p, _c.c_int, _c.c_double, _c.c_float, _c.c_size_t
_pp = _c.POINTER(_c.c_void_p)
_ip = _c.POINTER(_c.c_int)

# name -> (restype, argtypes).  Must list every symbol declared in include/adpst.h (tests/test_abi.py checks).
SIGNATURES = {
    "adpst_version": (_i, []),
    "adpst_last_error": (_c.c_char_p, []),
    "adpst_launch_count": (_c.c_ulonglong, []),
    "adpst_laplacian_create": (_i, [_i, _i, _i, _i, _d, _vp, _i, _i, _vp, _pp]),
    "adpst_laplacian_destroy": (None, [_vp]),
    "adpst_laplacian_matvec": (_i, [_vp, _vp, _vp, _d, _vp, _vp]),
    "adpst_laplacian_set_quadratic_window": (_i, [_vp, _i, _i]),
    "adpst_laplacian_set_kernel": (_i, [_vp, _i, _vp]),
    "adpst_laplacian_kernel": (_i, [_vp]),
    "adpst_laplacian_coefficients": (_i, [_vp, _vp, _vp, _vp]),
    "adpst_laplacian_nnz": (_c.c_int64, [_vp]),
    "adpst_laplacian_export_coo": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "adpst_adam_clip_step": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _f, _f, _f, _f, _vp]),
    "adpst_vgg_create": (_i, [_pp, _pp, _vp, _pp]),
    "adpst_vgg_destroy": (None, [_vp]),
    "adpst_vgg_conv_shape": (_i, [_i, _i, _i, _c.POINTER(_i), _c.POINTER(_i), _c.POINTER(_i)]),
    "adpst_vgg_pool_shape": (_i, [_i, _i, _i, _c.POINTER(_i), _c.POINTER(_i), _c.POINTER(_i)]),
    "adpst_vgg_forward": (_i, [_vp, _vp, _i, _i, _pp, _pp, _i, _vp]),
    "adpst_vgg_forward_range": (_i, [_vp, _vp, _i, _i, _pp, _pp, _i, _i, _ip, _ip, _vp]),
    "adpst_vgg_backward_range": (_i, [_vp, _i, _i, _pp, _pp, _i, _i, _vp, _vp, _vp, _vp, _ip, _ip, _vp]),
    "adpst_absmax_update": (_i, [_vp, _sz, _vp, _vp]),
    "adpst_vgg_grad_absmax": (_vp, [_vp, _i]),
    "adpst_halo_create": (_i, [_sz, _pp]),
    "adpst_halo_destroy": (None, [_vp]),
    "adpst_halo_ipc_handle_bytes": (_i, []),
    "adpst_halo_export": (_i, [_vp, _vp]),
    "adpst_halo_connect_ipc": (_i, [_vp, _i, _vp, _sz]),
    "adpst_halo_connect_local": (_i, [_vp, _i, _vp]),
    "adpst_halo_push": (_i, [_vp, _i, _sz, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "adpst_halo_pull": (_i, [_vp, _i, _sz, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "adpst_vgg_set_conv_path": (_i, [_vp, _i]),
    "adpst_absmax": (_i, [_vp, _sz, _vp, _vp]),
    "adpst_vgg_act_absmax": (_vp, [_vp, _i]),
    "adpst_vgg_conv_forward": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "adpst_vgg_conv_dgrad": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "adpst_vgg_backward": (_i, [_vp, _i, _i, _pp, _pp, _pp, _i, _vp, _vp, _vp, _vp]),
    "adpst_resize_bilinear": (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    "adpst_resize_bilinear_batch": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "adpst_gram_workspace_bytes": (_sz, [_i, _i, _i]),
    "adpst_gram_masked": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "adpst_style_layer_backward": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _d, _d, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp]),
    "adpst_style_tiles_bytes": (_sz, [_i]),
    "adpst_style_tiles": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "adpst_content_layer": (_i, [_vp, _vp, _sz, _d, _d, _vp, _vp, _i, _d, _i, _i, _i, _i, _vp]),
    "adpst_tv_loss": (_i, [_vp, _i, _i, _d, _d, _vp, _vp, _i, _i, _i, _vp]),
    "adpst_loss_finalize": (_i, [_vp, _d, _d, _d, _d, _vp, _vp]),
    "adpst_axpby": (_i, [_vp, _vp, _f, _vp, _f, _sz, _vp]),
}

_lib = None


def lib():
    """The loaded CDLL.  Raises if libadpst.so has not been built (python __graft_entry__.py / build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libadpst.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name, None)
            if fn is None:                      # reported by missing_symbols(); calling it raises AttributeError
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def missing_symbols():
    L = lib()
    return [n for n in SIGNATURES if not hasattr(L, n)]


def launch_count():
    return int(lib().adpst_launch_count())


def check(code):
    if code != OK:
        raise AdpstError(code, lib().adpst_last_error().decode("utf-8", "replace"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("automated-deep-photo-style-transfer_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def dtype_code(dt):
    import torch
    if dt == torch.float32:
        return F32
    if dt == torch.float64:
        return F64
    raise TypeError("unsupported dtype %s (float32 / float64 only)" % dt)
