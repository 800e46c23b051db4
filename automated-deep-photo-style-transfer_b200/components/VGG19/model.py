"""Drop-in for /root/reference/components/VGG19/model.py (vgg_layers :4-16, StyleContentModel :18-41).

Keras `applications.vgg19` is replaced by the VGG19 handle of libadpst (csrc/vgg_simt.cu, conv_tc.cu): the x255,
RGB->BGR, mean subtraction of `call` (:28-29) is folded into the first convolution's load.  There is no autodiff
tape; `backward(seeds)` plays the role of tape.gradient (style_transfer.py:341) for the frozen network.
"""
import ctypes
import os
import weakref

import numpy as np
import torch

from ... import _lib
from ...synth import CONV_LAYERS

LAYER_INDEX = {name: i for i, (name, _, _) in enumerate(CONV_LAYERS)}
POOL_AFTER = (1, 3, 7, 11)
BLOCKS = ((0, 1), (2, 3), (4, 7), (8, 11), (12, 12))      # conv index ranges of block1 .. block5
# Segments of a spatially tiled pass (tiled.py): the blocks, with block4 cut in two so that no segment has more than two
# convolutions at 1/8 resolution -- a halo of h columns supports n <= h / 2 convolutions between two exchanges.
SEGMENTS = ((0, 1), (2, 3), (4, 7), (8, 9), (10, 11), (12, 12))


def _load_weights(weights):
    """weights: dict name -> (kernel HWIO (3,3,Cin,Cout), bias (Cout,)), or a path to an .npz with
    '<name>/kernel' and '<name>/bias' entries.  The reference downloads ImageNet weights through Keras
    (model.py:7); there is no network here, so they must be supplied."""
    if weights is None:
        path = os.environ.get("ADPST_VGG19_WEIGHTS", os.path.join(os.path.dirname(__file__), "..", "..", "..", "weights",
                                                                 "vgg19_conv.npz"))
        if not os.path.exists(path):
            raise FileNotFoundError(
                "VGG19 weights not found (%s). Pass weights=<dict or .npz path> or set ADPST_VGG19_WEIGHTS; "
                "Keras' ImageNet download of the reference (VGG19/model.py:7) is not available offline." % path)
        weights = path
    if isinstance(weights, (str, os.PathLike)):
        z = np.load(weights)
        weights = {n: (z[n + "/kernel"], z[n + "/bias"]) for n, _, _ in CONV_LAYERS}
    out = []
    for name, cin, cout in CONV_LAYERS:
        k, b = weights[name]
        k, b = np.asarray(k, np.float32), np.asarray(b, np.float32)
        if k.shape != (3, 3, cin, cout) or b.shape != (cout,):
            raise ValueError("%s: expected kernel (3,3,%d,%d) and bias (%d,), got %s / %s" %
                             (name, cin, cout, cout, k.shape, b.shape))
        out.append((k, b))
    return out


class VGG19Handle:
    """Owns one adpst_vgg* (weights re-laid-out on the device)."""

    def __init__(self, weights=None, device=None):
        _lib.require_cuda()
        self.device = torch.device(device if device is not None else "cuda")
        kb = _load_weights(weights)
        ks = [torch.as_tensor(k).to(self.device).contiguous() for k, _ in kb]
        bs = [torch.as_tensor(b).to(self.device).contiguous() for _, b in kb]
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_vgg_create(_lib.ptr_array(ks), _lib.ptr_array(bs), _lib.stream_ptr(),
                                                   ctypes.byref(h)))
            torch.cuda.current_stream().synchronize()        # ks / bs may be freed after this
        self._h = h
        self.generation = 0              # forward passes run on this handle (its max|activation| slots describe the latest)
        if os.environ.get("ADPST_CONV_PATH", "").lower() == "simt":      # validation switch: exact-fp32 CUDA-core kernels
            self.set_conv_path("simt")

    def set_conv_path(self, path):
        """'tensor' (default: tcgen05 3xFP16 implicit GEMM) or 'simt' (exact float32 CUDA-core kernels, validation)."""
        _lib.check(_lib.lib().adpst_vgg_set_conv_path(self._h, {"tensor": 0, "simt": 1}[path]))

    def act_absmax_ptr(self, i):
        """Device address of the slot holding max|conv i output| of the latest forward pass."""
        return _lib.lib().adpst_vgg_act_absmax(self._h, i)

    def grad_absmax_ptr(self, i):
        """Device address of the max|dLoss/d(pre-activation of conv i)| word of the latest backward pass."""
        return _lib.lib().adpst_vgg_grad_absmax(self._h, i)

    def absmax_update(self, t, slot):
        """Raise the scale word at device address `slot` (act_absmax_ptr / grad_absmax_ptr) to max|t| if that is larger (t: data
        patched into the tensor after its producer ran, e.g. halo columns received from a neighbouring rank)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_absmax_update(_lib.ptr(t), t.numel(), ctypes.c_void_p(slot), _lib.stream_ptr()))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().adpst_vgg_destroy(h)
            except Exception:
                pass

    @staticmethod
    def conv_shape(i, H, W):
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().adpst_vgg_conv_shape(i, H, W, ctypes.byref(h), ctypes.byref(w), ctypes.byref(c)))
        return h.value, w.value, c.value

    @staticmethod
    def pool_shape(j, H, W):
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().adpst_vgg_pool_shape(j, H, W, ctypes.byref(h), ctypes.byref(w), ctypes.byref(c)))
        return h.value, w.value, c.value


class Activations:
    """Caller-owned activation set of one forward pass (conv outputs post-ReLU and pool outputs)."""

    def __init__(self, H, W, last, device, geom=None):
        """geom (strips of a tiled run, see adpst_vgg_forward_range): (widths of the five resolution levels, column offsets of
        the four pooled tensors); such tensors are zero-filled once, because the columns next to a pooled tensor are only
        ever written by the halo exchange."""
        self.H, self.W, self.last, self.geom = H, W, last, geom
        self.acts = [None] * _lib.VGG_NUM_CONV
        self.pools = [None] * _lib.VGG_NUM_POOL
        make = torch.empty if geom is None else torch.zeros
        for i in range(last + 1):
            h, w, c = VGG19Handle.conv_shape(i, H, W)
            if geom is not None:
                w = geom[0][sum(1 for p in POOL_AFTER if p < i)]
            if h < 1 or w < 1:
                raise ValueError("a %dx%d image is too small for VGG19 layer %s" % (H, W, CONV_LAYERS[i][0]))
            self.acts[i] = make(1, h, w, c, dtype=torch.float32, device=device)
        for j, after in enumerate(POOL_AFTER):
            if after < last:
                h, w, c = VGG19Handle.pool_shape(j, H, W)
                if geom is not None:
                    w = geom[0][j + 1]
                self.pools[j] = make(1, h, w, c, dtype=torch.float32, device=device)
        self.c_geom = (None, None) if geom is None else ((ctypes.c_int * 5)(*geom[0]), (ctypes.c_int * 4)(*geom[1]))


def vgg_layers(layer_names, shape=None, weights=None, device=None):
    """reference :4-16.  Returns the handle and the conv indices of the requested layers."""
    for n in layer_names:
        if n not in LAYER_INDEX:
            raise ValueError("unknown VGG19 layer %r (available: %s)" % (n, ", ".join(LAYER_INDEX)))
    return VGG19Handle(weights, device), [LAYER_INDEX[n] for n in layer_names]


class StyleContentModel:
    """reference :18-41.  `model(image)` -> {'content': {name: (1,h,w,C)}, 'style': {name: (1,h,w,C)}}, post-ReLU.

    Extensions (no tape here): the activations of the latest call stay referenced in `self.last` so that
    `backward(seeds)` can return d(sum_i <seed_i, layer_i>)/d(image)."""

    def __init__(self, content_layers, style_layers, shape=None, weights=None, device=None):
        self.vgg, idx = vgg_layers(list(content_layers) + list(style_layers), shape, weights, device)
        self.content_layers = list(content_layers)
        self.style_layers = list(style_layers)
        self.limit = len(content_layers)                                      # :24
        self.indices = idx
        self.last_index = max(idx)
        self.device = self.vgg.device
        self.last = None
        self._loop = None
        self._scratch = None

    def __call__(self, inputs, reuse=False):
        return self.call(inputs, reuse)

    def call(self, inputs, reuse=False):                                      # :27-41
        """reuse=True overwrites the activation buffers of the previous call (what train_step does every
        iteration); the default allocates fresh tensors, so earlier results (the targets) stay valid."""
        if inputs.dim() != 4 or inputs.shape[0] != 1 or inputs.shape[3] != 3:
            raise ValueError("expected an image of shape (1, H, W, 3), got %s" % (tuple(inputs.shape),))
        if inputs.dtype != torch.float32 or not inputs.is_cuda:
            raise TypeError("expected a float32 CUDA image")
        x = inputs.contiguous()
        H, W = int(x.shape[1]), int(x.shape[2])
        if reuse:
            # a private buffer set that only reuse=True calls ever write: results handed out by reuse=False calls
            # (the content / style targets) are never overwritten
            if self._loop is None or (self._loop.H, self._loop.W, self._loop.geom) != (H, W, None):
                self._loop = Activations(H, W, self.last_index, self.device)
            A = self._loop
        else:
            A = Activations(H, W, self.last_index, self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_vgg_forward(self.vgg._h, _lib.ptr(x), H, W, _lib.ptr_array(A.acts),
                                                    _lib.ptr_array(A.pools), self.last_index, _lib.stream_ptr()))
        self.last = A
        self.vgg.generation += 1
        names = self.content_layers + self.style_layers
        outs = [A.acts[i] for i in self.indices]
        for i, o in zip(self.indices, outs):                                  # see kernels.act_absmax_slot
            o._adpst_absmax = (weakref.ref(self.vgg), self.vgg.generation, i)
        content = {n: o for n, o in zip(names[:self.limit], outs[:self.limit])}
        style = {n: o for n, o in zip(names[self.limit:], outs[self.limit:])}
        return {"content": content, "style": style}

    # ---- spatially tiled runs (tiled.py): block by block, with a halo exchange on every tensor that crosses a pool --------
    def forward_blocks(self, inputs, exchange, reuse=True, overlap=None, geom=None):
        """Like call(), but the network runs one segment (SEGMENTS) at a time and `exchange(tensor, slot)` is called on the
        tensor that leaves a segment -- the pooled tensor, or conv 9's output between the two halves of block4 -- before the
        next segment reads it.  `exchange` overwrites the halo columns in place with the neighbours' data and raises the scale
        slot `slot` (device address of the tensor's max|.| word, see VGG19Handle.absmax_update) to the maximum of what arrived.
        overlap(last, outputs): optional; work that only needs the layers up to conv `last`, handed to exchange() as a third
        argument (a callable without arguments) to be enqueued while the halo columns are in flight.
        geom: (level widths, pool column offsets) of a strip with per-level halos (Activations)."""
        if inputs.dim() != 4 or inputs.shape[0] != 1 or inputs.shape[3] != 3 or inputs.dtype != torch.float32 or not inputs.is_cuda:
            raise TypeError("expected a float32 CUDA image of shape (1, H, W, 3)")
        x = inputs.contiguous()
        H, W = int(x.shape[1]), int(x.shape[2])
        if geom is not None:
            geom = (tuple(int(v) for v in geom[0]), tuple(int(v) for v in geom[1]))
        if reuse:
            if self._loop is None or (self._loop.H, self._loop.W, self._loop.geom) != (H, W, geom):
                self._loop = Activations(H, W, self.last_index, self.device, geom)
            A = self._loop
        else:
            A = Activations(H, W, self.last_index, self.device, geom)
        L = _lib.lib()
        self.last = A
        self.vgg.generation += 1
        names = self.content_layers + self.style_layers
        outs = [A.acts[i] for i in self.indices]
        for i, o in zip(self.indices, outs):
            o._adpst_absmax = (weakref.ref(self.vgg), self.vgg.generation, i)
        content = {n: o for n, o in zip(names[:self.limit], outs[:self.limit])}
        style = {n: o for n, o in zip(names[self.limit:], outs[self.limit:])}
        outputs = {"content": content, "style": style}
        for first, last in SEGMENTS:
            if first > self.last_index:
                break
            last = min(last, self.last_index)
            with torch.cuda.device(self.device):
                _lib.check(L.adpst_vgg_forward_range(self.vgg._h, _lib.ptr(x), H, W, _lib.ptr_array(A.acts),
                                                     _lib.ptr_array(A.pools), first, last, A.c_geom[0], A.c_geom[1],
                                                     _lib.stream_ptr()))
            if last == self.last_index:
                break
            out = A.pools[POOL_AFTER.index(last)] if last in POOL_AFTER else A.acts[last]
            # a pooled tensor shares the scale slot of the conv before it
            exchange(out, self.vgg.act_absmax_ptr(last), None if overlap is None else (lambda l=last: overlap(l, outputs)))
        return outputs

    def backward_blocks(self, seeds, exchange, out=None, overlap=None):
        """Like backward(), segment by segment from the top: the gradient that leaves a segment -- w.r.t. the pooled tensor
        below it, or w.r.t. conv 9's pre-activation between the two halves of block4 -- is handed to `exchange(tensor, slot)`
        (halo columns replaced by the owners' complete values; slot: None, or the scale word the consumer reads for that
        gradient) before the segment below uses it.
        overlap(first): optional.  The seed tensors may be filled lazily: overlap(first) is called (a) directly, with the first
        conv index of the segment that is about to run, and must make sure that the seeds of every layer >= first hold their
        values, and (b) through exchange()'s third argument with first = -1 while halo columns are in flight, where it may
        produce any seed that is still missing."""
        A = self.last
        if A is None:
            raise RuntimeError("backward_blocks() needs a preceding forward call")
        arr = [None] * _lib.VGG_NUM_CONV
        for n, t in seeds.items():
            i = LAYER_INDEX[n]
            if t.shape != A.acts[i].shape or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("seed for %s must be a contiguous float32 tensor of shape %s" % (n, tuple(A.acts[i].shape)))
            arr[i] = t
        top = max(i for i, t in enumerate(arr) if t is not None)
        if self._scratch is None or self._scratch[0].numel() < A.acts[0].numel():
            self._scratch = (torch.empty(A.acts[0].numel(), dtype=torch.float32, device=self.device),
                             torch.empty(A.acts[0].numel(), dtype=torch.float32, device=self.device))
        if out is None:
            out = torch.empty(1, A.H, A.W, 3, dtype=torch.float32, device=self.device)
        if self._dseg is None or self._dseg_shape != (A.H, A.W, A.geom):
            # the gradient that leaves segment s (entering at conv `first`): shaped like that segment's input
            self._dseg = {}
            for first, _ in SEGMENTS[1:]:
                src = A.pools[POOL_AFTER.index(first - 1)] if (first - 1) in POOL_AFTER else A.acts[first - 1]
                if src is not None:
                    self._dseg[first] = torch.empty_like(src)
            self._dseg_shape = (A.H, A.W, A.geom)
        L = _lib.lib()
        grad_in = None
        for first, last in reversed(SEGMENTS):
            if first > top:
                continue
            last = min(last, top)
            target = out if first == 0 else self._dseg[first]
            if overlap is not None:
                overlap(first)
            with torch.cuda.device(self.device):
                _lib.check(L.adpst_vgg_backward_range(self.vgg._h, A.H, A.W, _lib.ptr_array(A.acts), _lib.ptr_array(arr), first,
                                                      last, _lib.ptr(grad_in), _lib.ptr(self._scratch[0]),
                                                      _lib.ptr(self._scratch[1]), _lib.ptr(target), A.c_geom[0], A.c_geom[1],
                                                      _lib.stream_ptr()))
            if first > 0:
                # mid-block: the consumer reads the scale slot of conv first-1's gradient; below a pool the un-pooling kernel
                # measures its own output
                exchange(target, None if (first - 1) in POOL_AFTER else self.vgg.grad_absmax_ptr(first - 1),
                         None if overlap is None else (lambda: overlap(-1)))
                grad_in = target
        return out

    _dseg = None
    _dseg_shape = None

    def backward(self, seeds, out=None):
        """seeds: dict layer name -> dLoss/d(layer output) (1,h,w,C) float32.  Returns dLoss/d(image) (1,H,W,3)."""
        A = self.last
        if A is None:
            raise RuntimeError("backward() needs a preceding forward call")
        arr = [None] * _lib.VGG_NUM_CONV
        for n, t in seeds.items():
            i = LAYER_INDEX[n]
            if t.shape != A.acts[i].shape or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("seed for %s must be a contiguous float32 tensor of shape %s" %
                                 (n, tuple(A.acts[i].shape)))
            arr[i] = t
        top = max(i for i, t in enumerate(arr) if t is not None)
        if self._scratch is None or self._scratch[0].numel() < A.acts[0].numel():
            self._scratch = (torch.empty(A.acts[0].numel(), dtype=torch.float32, device=self.device),
                             torch.empty(A.acts[0].numel(), dtype=torch.float32, device=self.device))
        if out is None:
            out = torch.empty(1, A.H, A.W, 3, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_vgg_backward(self.vgg._h, A.H, A.W, _lib.ptr_array(A.acts), _lib.ptr_array(A.pools),
                                                     _lib.ptr_array(arr), top, _lib.ptr(self._scratch[0]),
                                                     _lib.ptr(self._scratch[1]), _lib.ptr(out), _lib.stream_ptr()))
        return out
