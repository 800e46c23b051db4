"""Thin, typed Python wrappers over the C-ABI entry points that are not tied to a handle.
Every function launches hand-written CUDA from libadpst.so on torch's current stream; none has a fallback."""
import ctypes
import functools
import os

import torch

from . import _lib


def _on_tensor_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current: the launch then goes to THAT device's current
    stream (an object built with device='cuda:N' may be driven while another device is current)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        for a in list(args) + list(kw.values()):
            t = a.m if isinstance(a, AdamState) else a
            if isinstance(t, torch.Tensor) and t.is_cuda:
                with torch.cuda.device(t.device):
                    return fn(*args, **kw)
        return fn(*args, **kw)
    return wrapper


def _f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError("%s must be a contiguous float32 CUDA tensor" % name)
    return t


class AdamState:
    """Adam slots for one image variable (style_transfer.py:321-326): m, v and the device-side step counter."""

    def __init__(self, like):
        self.m = torch.zeros_like(like)
        self.v = torch.zeros_like(like)
        self.state = torch.zeros(2, dtype=torch.int32, device=like.device)     # {t, ticket}

    @property
    def step(self):
        return int(self.state[0])


@_on_tensor_device
def adam_clip_step(x, grad, st, lr=0.1, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """In place: TF-flavour Adam update of x followed by clip to [0,1] (style_transfer.py:342-343)."""
    _f32(x, "x"); _f32(grad, "grad")
    if grad.numel() != x.numel():
        raise ValueError("grad and x differ in size")
    _lib.check(_lib.lib().adpst_adam_clip_step(_lib.ptr(x), _lib.ptr(grad), _lib.ptr(st.m), _lib.ptr(st.v), x.numel(),
                                               _lib.ptr(st.state), lr, beta1, beta2, epsilon, _lib.stream_ptr()))
    return x


@_on_tensor_device
def resize_bilinear(mask_hw, size):
    """tf.image.resize(mask, size) for one (H,W) float32 plane (loss.py:112-113)."""
    _f32(mask_hw, "mask")
    Hs, Ws = mask_hw.shape
    Hd, Wd = int(size[0]), int(size[1])
    out = torch.empty(Hd, Wd, dtype=torch.float32, device=mask_hw.device)
    _lib.check(_lib.lib().adpst_resize_bilinear(_lib.ptr(mask_hw), Hs, Ws, _lib.ptr(out), Hd, Wd, _lib.stream_ptr()))
    return out


@_on_tensor_device
def resize_bilinear_batch(planes, size):
    """tf.image.resize for a stack of planes (n,H,W) -> (n,h,w): all class masks of a layer in one launch (loss.py:112-117)."""
    _f32(planes, "planes")
    n, Hs, Ws = planes.shape
    Hd, Wd = int(size[0]), int(size[1])
    out = torch.empty(n, Hd, Wd, dtype=torch.float32, device=planes.device)
    _lib.check(_lib.lib().adpst_resize_bilinear_batch(_lib.ptr(planes), n, Hs, Ws, _lib.ptr(out), Hd, Wd, _lib.stream_ptr()))
    return out


@_on_tensor_device
def content_layer(target, output, loss_scale, grad_scale, loss_acc, d_out=None, accumulate=False, n_norm=0.0, own_cols=None):
    """loss_acc (float64[1]) += loss_scale*mean((t-o)^2); d_out (=|+=) grad_scale*2(o-t)/n  (loss.py:90-92).
    Spatially tiled runs: n_norm = element count of the whole layer, own_cols = (lo, hi) columns of this (1,h,w,C) tile
    that count towards the scalar."""
    _f32(target, "target"); _f32(output, "output")
    if target.shape != output.shape:
        raise ValueError("content target %s and output %s differ in shape" % (tuple(target.shape), tuple(output.shape)))
    _lib.check(_lib.lib().adpst_content_layer(_lib.ptr(target), _lib.ptr(output), output.numel(), float(loss_scale),
                                              float(grad_scale), _lib.ptr(loss_acc), _lib.ptr(d_out), int(bool(accumulate)),
                                              float(n_norm), *((int(output.shape[2]), int(output.shape[3]), int(own_cols[0]),
                                                                int(own_cols[1])) if own_cols is not None else (0, 0, 0, 0)),
                                              _lib.stream_ptr()))


@_on_tensor_device
def axpby(out, a, alpha, b=None, beta=0.0):
    _f32(out, "out"); _f32(a, "a")
    _lib.check(_lib.lib().adpst_axpby(_lib.ptr(out), _lib.ptr(a), float(alpha), _lib.ptr(b), float(beta), out.numel(),
                                      _lib.stream_ptr()))
    return out


def gram_workspace(HW, C, K, device):
    n = int(_lib.lib().adpst_gram_workspace_bytes(HW, C, K))
    return torch.empty(max(n, 16), dtype=torch.uint8, device=device)


PATCH_H, PATCH_W = 2, 16        # pixel patch of one pipeline stage of the tensor-core Gram kernel (csrc/gram_tc.cu)


def gram_patch_lists(masks, h, w, K, device):
    """Patches (2 x 16 pixels, row-major patch index) on which each class mask is non-zero, grouped by class.
    masks: (K, h*w) float32 or None.  Returns (patch_ids int32, patch_off int32[K+1]).  Host-side set-up (the masks
    are constant over the optimisation), plain tensor plumbing."""
    ph, pw = -(-h // PATCH_H), -(-w // PATCH_W)
    if masks is None:
        active = torch.ones(1, ph * pw, dtype=torch.bool, device=device)
    else:
        m = torch.zeros(K, ph * PATCH_H, pw * PATCH_W, dtype=torch.float32, device=device)
        m[:, :h, :w] = masks.reshape(K, h, w)
        active = (m.reshape(K, ph, PATCH_H, pw, PATCH_W) != 0).any(dim=4).any(dim=2).reshape(K, ph * pw)
    ids = active.nonzero()[:, 1].to(torch.int32).contiguous()
    off = torch.zeros(active.shape[0] + 1, dtype=torch.int32, device=device)
    off[1:] = active.sum(1).cumsum(0).to(torch.int32)
    if ids.numel() == 0:
        ids = torch.zeros(1, dtype=torch.int32, device=device)
    return ids, off


@_on_tensor_device
def absmax_slot(t):
    """A 1-element int32 tensor holding the float32 bit pattern of max|t| (the scale slot the tensor-core kernels read)."""
    slot = torch.zeros(1, dtype=torch.int32, device=t.device)
    _lib.check(_lib.lib().adpst_absmax(_lib.ptr(t), t.numel(), _lib.ptr(slot), _lib.stream_ptr()))
    return slot


# validation switch, like ADPST_CONV_PATH for the convolutions: "simt" routes the Gram and style-gradient kernels to the exact-float32
# CUDA-core kernels
_DEFAULT_PATH = "simt" if os.environ.get("ADPST_LOSS_PATH", "").lower() == "simt" else "tensor"


@_on_tensor_device
def gram_masked(F, masks, K, workspace=None, patches=None, path=None, out=None, f_absmax=None, masks_absmax=None):
    """F: (h,w,C) float32 feature map ((HW,C) is taken as h = HW, w = 1); masks: (K,h*w) float32 or None.
    Returns (K,C,C) float32  (loss.py:96-102).  `patches` = gram_patch_lists(...) enables the tcgen05 kernel."""
    path = path or _DEFAULT_PATH
    _f32(F, "F")
    if F.dim() == 2:
        F = F.reshape(F.shape[0], 1, F.shape[1])
    h, w, C = F.shape
    if masks is not None:
        _f32(masks, "masks")
        if tuple(masks.shape) != (K, h * w):
            raise ValueError("masks must have shape (K, h*w)")
    elif K != 1:
        raise ValueError("K must be 1 without masks")
    ws = workspace if workspace is not None else gram_workspace(h * w, C, K, F.device)
    G = out if out is not None else torch.empty(K, C, C, dtype=torch.float32, device=F.device)
    ids, off = patches if patches is not None else (None, None)
    _lib.check(_lib.lib().adpst_gram_masked(_lib.ptr(F), h, w, C, _lib.ptr(masks), K, _lib.ptr(ids), _lib.ptr(off), _lib.ptr(G),
                                            {"tensor": 0, "simt": 1}[path], ctypes.c_void_p(f_absmax or 0),
                                            _lib.ptr(masks_absmax), _lib.ptr(ws), _lib.stream_ptr()))
    return G


def act_absmax_slot(t):
    """Device address of the slot holding max|t| if `t` is an output of the latest forward pass of its VGG19 handle
    (components/VGG19/model.py tags its outputs), else None: the consumer then measures max|t| itself."""
    tag = getattr(t, "_adpst_absmax", None)
    if tag is None:
        return None
    handle, generation, index = tag
    handle = handle()
    if handle is None or handle.generation != generation:
        return None
    return handle.act_absmax_ptr(index)


def style_tiles(masks, K, h, w, device):
    """Set-up for style_layer_backward: classes present in each 8x16-pixel tile of the (constant) masks."""
    if K > 32:
        return None                      # the tensor-core style gradient handles at most 32 classes per launch
    out = torch.empty(int(_lib.lib().adpst_style_tiles_bytes(h * w)), dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().adpst_style_tiles(_lib.ptr(masks), K, h, w, _lib.ptr(out), _lib.stream_ptr()))
    return out


@_on_tensor_device
def style_layer_backward(F, masks, K, G, A, loss_scale, grad_scale, loss_acc, dF, accumulate=False, workspace=None,
                         path=None, hw_norm=0.0, f_absmax=None, tiles=None):
    """One layer of loss.py:104-137: accumulates the loss value and writes/adds its gradient w.r.t. F.
    F: (h, w, C) feature map (a (HW, C) matrix is treated as h = HW, w = 1... use the 3-D form for 2-D tiling).
    f_absmax: device address of a slot holding max|F| (act_absmax_slot), or None to have it measured."""
    path = path or _DEFAULT_PATH
    _f32(F, "F"); _f32(G, "G"); _f32(A, "A")
    if F.dim() == 2:
        F = F.reshape(F.shape[0], 1, F.shape[1])
    h, w, C = F.shape
    ws = workspace if workspace is not None else gram_workspace(h * w, C, K, F.device)
    _lib.check(_lib.lib().adpst_style_layer_backward(_lib.ptr(F), h, w, C, _lib.ptr(masks), K, _lib.ptr(G), _lib.ptr(A),
                                                     float(loss_scale), float(grad_scale), _lib.ptr(loss_acc), _lib.ptr(dF),
                                                     int(bool(accumulate)), {"tensor": 0, "simt": 1}[path], float(hw_norm),
                                                     ctypes.c_void_p(f_absmax or 0), _lib.ptr(tiles), _lib.ptr(ws),
                                                     _lib.stream_ptr()))


@_on_tensor_device
def loss_finalize(acc, w_content, w_style, w_photo, out, w_tv=0.0):
    """acc: float64[4] {content, style, photo, tv}; out: float32[6] {content, style, nima, photo, total, tv}  (loss.py:72-76;
    the tv term is an extension and enters the total only if w_tv > 0)."""
    if acc.numel() < 4 or out.numel() < 6:
        raise ValueError("loss_finalize: acc needs 4 float64 entries and out 6 float32 entries")
    _lib.check(_lib.lib().adpst_loss_finalize(_lib.ptr(acc), float(w_content), float(w_style), float(w_photo), float(w_tv),
                                              _lib.ptr(out), _lib.stream_ptr()))
    return out


@_on_tensor_device
def tv_loss(image, loss_scale, grad_scale, loss_acc, d_image=None, accumulate=False, own_cols=None):
    """EXTENSION (not in the reference): tf.image.total_variation of a (1,H,W,3) float32 image.
    loss_acc (float64[1], may be None) += loss_scale * TV;  d_image (=|+=) grad_scale * dTV/dx."""
    _f32(image, "image")
    if image.dim() != 4 or image.shape[0] != 1 or image.shape[3] != 3:
        raise ValueError("expected an image of shape (1, H, W, 3), got %s" % (tuple(image.shape),))
    if d_image is not None:
        _f32(d_image, "d_image")
        if d_image.numel() != image.numel():
            raise ValueError("d_image and image differ in size")
    lo, hi = (0, 0) if own_cols is None else (int(own_cols[0]), int(own_cols[1]))
    _lib.check(_lib.lib().adpst_tv_loss(_lib.ptr(image), int(image.shape[1]), int(image.shape[2]), float(loss_scale),
                                        float(grad_scale), _lib.ptr(loss_acc), _lib.ptr(d_image), int(bool(accumulate)),
                                        lo, hi, _lib.stream_ptr()))
