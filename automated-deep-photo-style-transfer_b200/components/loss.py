"""Drop-in for /root/reference/components/loss.py (class Loss :6-161, dict_zip :163-165).

Same constructor and call signature, same dictionary keys (including the reference's spelling of
'Photorealism regualarization').  Differences forced by the platform, all explicit:
  * there is no autodiff tape: evaluating the loss also writes the gradient seeds (d total / d layer output) and the
    photorealism gradient, which `gradient(extractor)` turns into d total / d image;
  * the NIMA term (loss.py:141-153, InceptionResNetV2 with weights that are not in the tree) is out of scope:
    `nima_weight` must be 0 and 'NIMA loss' is reported as 0;
  * the style Grams of the style target are constant and are computed once here, not every call (loss.py:130);
  * `matting` selects the Laplacian variant ('v2' as hard-wired in the reference, loss.py:4, or 'v3');
  * `args.tv_weight` (optional attribute, default 0) adds a total-variation term -- an EXTENSION with no counterpart in
    the reference.  With the default the dictionary, the total and the gradient are exactly the reference's.
All arithmetic happens in libadpst kernels; the returned values are views of one float32 device vector.
"""
import warnings

import torch

from .. import kernels
from .matting_v2 import MattingLaplacian as MattingLaplacianV2
from .matting_v3 import MattingLaplacian as MattingLaplacianV3


class Loss:
    r"""Loss functions are computed within this class (reference loss.py:6-9)."""

    def __init__(self, content_target, style_target, args, content_masks=None, style_masks=None, matting="v2",
                 tile=None, style_tile=None):
        """tile / style_tile (extensions, see tiled.py): `Tile` descriptions when the transfer / style image handed to
        this object is one column strip (with halo) of a larger image that is spread over several GPUs."""
        self.tile, self.style_tile = tile, style_tile
        self.content_target = content_target
        self.style_target = style_target
        self.content_masks = content_masks
        self.style_masks = style_masks

        self.loss_names = {                                                    # loss.py:16-21
            'content': 'Content loss',
            'style':   'Style loss',
            'nima':    'NIMA loss',
            'photo':   'Photorealism regualarization'
        }
        self.loss_weights = {                                                  # loss.py:28-33
            'content': float(args.content_weight),
            'style':   float(args.style_weight),
            'nima':    float(getattr(args, 'nima_weight', 0.0)),
            'photo':   float(args.regularization_weight)
        }
        if self.loss_weights['nima'] != 0.0:
            raise NotImplementedError("the NIMA term (loss.py:141-153) is not part of this hot path: nima_weight must "
                                      "be 0 (got %g)" % self.loss_weights['nima'])
        if not Loss._nima_note_shown:
            Loss._nima_note_shown = True
            warnings.warn("automated-deep-photo-style-transfer_b200: the reference adds w_nima * (10 - NIMA score) with a default "
                          "nima_weight of 1e5 (loss.py:64-72); that term is not built here, 'NIMA loss' is reported as 0 and "
                          "'Total loss' / the gradient match a reference run with --nima_weight 0 only", stacklevel=2)
        # EXTENSION (not in the reference): total-variation weight, tf.image.total_variation semantics; 0 = off
        self.tv_weight = float(getattr(args, 'tv_weight', 0.0) or 0.0)
        self.tv_name = 'Total variation loss'
        if (content_masks is None) != (style_masks is None):
            pass        # reference: masks are used only if both are given (loss.py:110); otherwise all-ones
        if content_masks is not None and style_masks is not None and len(content_masks) != len(style_masks):
            raise ValueError("content and style masks differ in count (%d vs %d)" % (len(content_masks), len(style_masks)))
        if matting not in ("v2", "v3"):
            raise ValueError("matting must be 'v2' or 'v3'")
        self.matting_variant = matting
        self.matting_params = {                                                # loss.py:38-41
            'epsilon': args.matting_epsilon,
            'window_radius': args.matting_window_radius,
        }
        self.matting_laplacian = None

        any_t = next(iter(content_target.values())) if len(content_target) else next(iter(style_target.values()))
        self.device = any_t.device
        self._acc = torch.zeros(4, dtype=torch.float64, device=self.device)    # content, style, photo, tv (unweighted)
        self._out = torch.zeros(6, dtype=torch.float32, device=self.device)    # content, style, nima, photo, total, tv
        self._layer_cache = {}      # style layer name -> dict(masks, K, A, ws, seed)
        self._content_seeds = {}
        self._photo_grad = None
        self._seeds = None
        self._seed_jobs = []
        self._part = None

    def initialize_matting_laplacian(self, image):                            # loss.py:45-46
        """image: (H,W,3) float64 (as in style_transfer.py:315) or float32.  If every value is exactly representable
        in float32 (always true for the script: the image was decoded to float32), HBM traffic is float32 and the
        stencil arithmetic float64; otherwise everything is float64."""
        img = torch.as_tensor(image)
        if not img.is_cuda:
            img = img.to(self.device)
        exact = img.dtype == torch.float32 or bool((img.to(torch.float32).to(img.dtype) == img).all())
        cls = MattingLaplacianV2 if self.matting_variant == "v2" else MattingLaplacianV3
        self.matting_laplacian = cls(img, storage_dtype=torch.float32 if exact else torch.float64,
                                     compute_dtype=torch.float64, **self.matting_params)
        if self.tile is not None:        # the scalar x^T L x counts this rank's own columns only
            self.matting_laplacian._op.set_quadratic_window(*self.tile.own_cols(int(img.shape[1])))

    def __call__(self, image, outputs):                                       # loss.py:48-49
        return self.compute_loss(image, outputs)

    # ------------------------------------------------------------------------------------------
    def _style_layer_state(self, name, target, output):
        st = self._layer_cache.get(name)
        if st is not None and st["shape"] == tuple(output.shape):
            return st
        _, h, w, C = output.shape
        _, hs, ws_, Cs = target.shape
        if C != Cs:
            raise ValueError("%s: target has %d channels, output %d" % (name, Cs, C))
        exist = self.content_masks is not None and self.style_masks is not None     # loss.py:110
        if exist:
            K = len(self.content_masks)
            if self._mask_planes is None:       # the full-resolution masks go to the device once, as two (K,H,W) stacks
                self._mask_planes = (torch.stack([_plane(m, self.device) for m in self.content_masks]).contiguous(),
                                     torch.stack([_plane(m, self.device) for m in self.style_masks]).contiguous())
            cplanes, splanes = self._mask_planes
            if self.tile is not None:            # a strip keeps a different number of halo columns on every resolution level
                a, b = self.tile.mask_cols(w)
                cplanes = cplanes[:, :, a:b].contiguous()
            if self.style_tile is not None:
                a, b = self.style_tile.mask_cols(ws_)
                splanes = splanes[:, :, a:b].contiguous()
            cm = kernels.resize_bilinear_batch(cplanes, (h, w)).reshape(K, h * w)                      # loss.py:112-117
            sm = kernels.resize_bilinear_batch(splanes, (hs, ws_)).reshape(K, hs * ws_)
        else:
            K, cm, sm = 1, None, None                                                # loss.py:119-120
        # spatial tiling: the Gram partial of this rank counts only its own columns (mask x column indicator; exact,
        # because strip boundaries are multiples of 16 px and bilinear half-pixel resizing never straddles them), while the
        # gradient uses the full local masks and the normaliser the pixel count of the whole image
        cm_own = self.tile.own_masks(cm, K, h, w, self.device) if self.tile is not None else cm
        sm_own = self.style_tile.own_masks(sm, K, hs, ws_, self.device) if self.style_tile is not None else sm
        ws = kernels.gram_workspace(max(h * w, hs * ws_), C, K, self.device)
        A = kernels.gram_masked(target.reshape(hs, ws_, C), sm_own, K, ws,
                                patches=kernels.gram_patch_lists(sm_own, hs, ws_, K, self.device))   # constant style Grams
        st = {"shape": tuple(output.shape), "K": K, "masks": cm, "own_masks": cm_own, "A": A, "ws": ws,
              "seed": torch.empty_like(output), "patches": kernels.gram_patch_lists(cm_own, h, w, K, self.device),
              "tiles": kernels.style_tiles(cm, K, h, w, self.device),
              "own_masks_absmax": kernels.absmax_slot(cm_own) if cm_own is not None else None,
              "G": torch.empty(K, C, C, dtype=torch.float32, device=self.device),
              "hw_norm": float(h * self.tile.global_cols(w)) if self.tile is not None else 0.0}
        self._layer_cache[name] = st
        return st

    def compute_loss(self, image, outputs):                                   # loss.py:53-78
        self.forward_partials(image, outputs)
        return self.finish()

    def style_targets_partial(self):
        """Spatially tiled runs: the style Grams A computed at set-up are per-rank partials; sum them over the ranks once."""
        return [st["A"] for st in self._layer_cache.values()]

    def prepare(self, outputs):
        """Build the per-layer state (masks, style Grams, buffers) for these output shapes without evaluating anything."""
        for name, target in self.style_target.items():
            self._style_layer_state(name, target, outputs['style'][name])

    def forward_partials(self, image, outputs):
        """First half of compute_loss: everything that is local to this image (or image strip).  Returns the tensors that
        must be summed over the ranks of a spatially tiled run before finish(): the per-layer Gram partials and the
        float64 accumulator {content, (unused), photo}.  A single-device run just calls finish() afterwards."""
        self.begin_partials(image)
        return self.end_partials(outputs)

    # The pieces of forward_partials, so that a tiled run can enqueue each one as soon as its inputs exist (while halo columns
    # of the next network segment are in flight): begin_partials, then partial_layers(outputs, upto) / partial_image() in any
    # order, then end_partials (which runs whatever is still missing).
    def begin_partials(self, image):
        self._acc.zero_()
        self._photo_grad = None
        self._part = {"image": image, "seeds": {}, "G": {}, "image_done": False}

    def partial_layers(self, outputs, upto):
        """Content terms and Gram partials of every layer whose conv index is <= upto and that has not been evaluated yet."""
        from .VGG19.model import LAYER_INDEX
        P, wts = self._part, self.loss_weights
        n_args = 2.0                                                          # len(args) in iter_on_layers, loss.py:85
        for name, target in self.content_target.items():                     # loss.py:59, :90-92
            if name in P["seeds"] or LAYER_INDEX.get(name, 0) > upto:     # names outside VGG19: no ordering
                continue
            out = outputs['content'][name]
            seed = self._content_seeds.get(name)
            if seed is None or seed.shape != out.shape:
                seed = self._content_seeds[name] = torch.empty_like(out)
            if self.tile is None:
                kernels.content_layer(target, out, 1.0 / n_args, wts['content'] / n_args, self._acc[0:1], seed)
            else:
                _, h, w, C = out.shape
                kernels.content_layer(target, out, 1.0 / n_args, wts['content'] / n_args, self._acc[0:1], seed,
                                      n_norm=float(h) * self.tile.global_cols(w) * C, own_cols=self.tile.own_cols(w))
            P["seeds"][name] = seed
        for name, target in self.style_target.items():                       # loss.py:62, :96-102
            if name in P["G"] or LAYER_INDEX.get(name, 0) > upto:
                continue
            out = outputs['style'][name]
            st = self._style_layer_state(name, target, out)
            _, h, w, C = out.shape
            P["G"][name] = kernels.gram_masked(out.reshape(h, w, C), st["own_masks"], st["K"], st["ws"], patches=st["patches"],
                                               out=st["G"], f_absmax=kernels.act_absmax_slot(out),
                                               masks_absmax=st["own_masks_absmax"])

    def partial_image(self):
        """The terms that read the image only: photorealism regulariser and (extension) total variation."""
        P, wts = self._part, self.loss_weights
        if P["image_done"]:
            return
        P["image_done"] = True
        image = P["image"]
        if wts['photo'] > 0:                                                  # loss.py:67-69, :157-161
            if self.matting_laplacian is None:
                raise RuntimeError("regularization_weight > 0 but initialize_matting_laplacian() was not called")
            self._photo_grad = self.calculate_photorealism_regularization(image, _with_gradient=True)
        if self.tv_weight > 0:                                                # extension: value -> _acc[3], gradient -> added
            own = None if self.tile is None else self.tile.own_cols(int(image.shape[2]))
            if self._photo_grad is None:
                if self._tv_grad_buf is None or self._tv_grad_buf.shape != image.shape:
                    self._tv_grad_buf = torch.empty_like(image)
                kernels.tv_loss(image, 1.0, self.tv_weight, self._acc[3:4], self._tv_grad_buf, accumulate=False, own_cols=own)
                self._photo_grad = self._tv_grad_buf
            else:
                kernels.tv_loss(image, 1.0, self.tv_weight, self._acc[3:4], self._photo_grad.reshape(image.shape),
                                accumulate=True, own_cols=own)

    def end_partials(self, outputs):
        self.partial_layers(outputs, 1 << 30)
        self.partial_image()
        P = self._part
        self._pending = (outputs, P["seeds"])
        return [P["G"][name] for name in self.style_target] + [self._acc]

    def finish(self, lazy=False):
        """Second half of compute_loss: style loss and its gradient seeds from the (global) Grams, weighted total.
        lazy (tiled.py): nothing is evaluated here; every seed tensor is handed out and filled by seed_layers(), the loss
        dictionary comes from finalize() once all of them have run."""
        from .VGG19.model import LAYER_INDEX
        outputs, seeds = self._pending
        self._acc[1:2].zero_()       # a cross-rank sum of the accumulator must not multiply the (global) style term
        self._seed_jobs = []
        for name in self.style_target:                                        # loss.py:104-137
            shared = name in seeds                                            # a layer can be both content and style
            dF = seeds[name] if shared else self._layer_cache[name]["seed"]
            self._seed_jobs.append((LAYER_INDEX.get(name, 0), name, dF, shared))
            seeds[name] = dF
        self._seed_jobs.sort()                                                # seed_layers() pops the deepest layer first
        self._seeds = seeds
        if lazy:
            return None
        self.seed_layers(0)
        return self.finalize()

    def seed_layers(self, down_to, at_most=None):
        """Evaluate the style term and gradient seed of the pending layers with conv index >= down_to, deepest first (at_most:
        stop after that many)."""
        outputs, _ = self._pending
        wts, n_args, done = self.loss_weights, 2.0, 0
        while self._seed_jobs and self._seed_jobs[-1][0] >= down_to and (at_most is None or done < at_most):
            _, name, dF, shared = self._seed_jobs.pop()
            out = outputs['style'][name]
            st = self._layer_cache[name]
            _, h, w, C = out.shape
            kernels.style_layer_backward(out.reshape(h, w, C), st["masks"], st["K"], st["G"], st["A"], 1.0 / n_args,
                                         wts['style'] / n_args, self._acc[1:2], dF.reshape(h * w, C), accumulate=shared,
                                         workspace=st["ws"], hw_norm=st["hw_norm"], f_absmax=kernels.act_absmax_slot(out),
                                         tiles=st["tiles"])
            done += 1

    def finalize(self):
        if self._seed_jobs:
            raise RuntimeError("finalize() before every style layer was evaluated")
        wts = self.loss_weights
        kernels.loss_finalize(self._acc, wts['content'], wts['style'], wts['photo'], self._out, w_tv=self.tv_weight)   # loss.py:72
        loss_dict = {self.loss_names['content']: self._out[0], self.loss_names['style']: self._out[1],
                     self.loss_names['nima']: self._out[2]}
        if wts['photo'] > 0:
            loss_dict[self.loss_names['photo']] = self._out[3]
        if self.tv_weight > 0:
            loss_dict[self.tv_name] = self._out[5]
        loss_dict['Total loss'] = self._out[4]                                # loss.py:76
        return loss_dict

    def gradient(self, extractor, out=None, backward=None):
        """d(Total loss)/d(image) for the latest compute_loss call: the role of tape.gradient in
        style_transfer.py:341.  `extractor` is the StyleContentModel whose outputs were passed in.
        backward (extension, tiled.py): callable (seeds, out) -> gradient that replaces extractor.backward."""
        if self._seeds is None:
            raise RuntimeError("gradient() needs a preceding compute_loss() call")
        g = extractor.backward(self._seeds, out=out) if backward is None else backward(self._seeds, out)
        if self._photo_grad is not None:
            kernels.axpby(g, g, 1.0, self._photo_grad.reshape(g.shape), 1.0)
        return g

    # ------------------------------------------------------------------------------------------
    # the reference's static helpers, evaluated on the GPU
    @staticmethod
    def calculate_layer_content_loss(target, output):                        # loss.py:90-92
        acc = torch.zeros(1, dtype=torch.float64, device=output.device)
        kernels.content_layer(target.contiguous(), output.contiguous(), 1.0, 0.0, acc, None)
        return acc[0].to(output.dtype)

    @staticmethod
    def calculate_gram_matrix(convolution_layer, mask):                      # loss.py:96-102
        _, h, w, C = convolution_layer.shape
        F = convolution_layer.reshape(h, w, C).contiguous()
        m = None if mask is None else mask.to(torch.float32).reshape(1, -1).contiguous()
        return kernels.gram_masked(F, m, 1, patches=kernels.gram_patch_lists(m, h, w, 1, F.device))[0]

    def calculate_layer_style_loss(self, target, output):                    # loss.py:104-137
        st = self._style_layer_state("<adhoc %s>" % (tuple(output.shape),), target, output)
        _, h, w, C = output.shape
        F = output.reshape(h, w, C).contiguous()
        G = kernels.gram_masked(F, st["masks"], st["K"], st["ws"], patches=st["patches"])
        acc = torch.zeros(1, dtype=torch.float64, device=output.device)
        kernels.style_layer_backward(F, st["masks"], st["K"], G, st["A"], 1.0, 0.0, acc, None, workspace=st["ws"])
        return acc[0].to(output.dtype)

    def calculate_photorealism_regularization(self, image, _with_gradient=False):   # loss.py:157-161
        lap = self.matting_laplacian
        HW = lap.shape[-1]
        p = image.reshape(HW, -1)
        if not _with_gradient:
            _, q = lap.quadratic_form(p, want_y=False)
            return q.to(image.dtype)
        if self._photo_grad_buf is None or self._photo_grad_buf.shape != p.shape or self._photo_grad_buf.dtype != lap._op.storage_dtype:
            self._photo_grad_buf = torch.empty(HW, 3, dtype=lap._op.storage_dtype, device=image.device)
        # y = 2 w_p L x is the gradient of w_p x^T L x (L symmetric); x^T L x lands in the float64 accumulator
        y, _ = lap.quadratic_form(p, want_y=True, y_scale=2.0 * self.loss_weights['photo'], out=self._photo_grad_buf,
                                  quad_out=self._acc[2:3])
        return y if y.dtype == torch.float32 else y.to(torch.float32)

    _photo_grad_buf = None
    _tv_grad_buf = None
    _mask_planes = None
    _nima_note_shown = False

    @staticmethod
    def calculate_total_variation(image):
        """EXTENSION: tf.image.total_variation(image)[0] of a (1,H,W,3) float32 CUDA image."""
        acc = torch.zeros(1, dtype=torch.float64, device=image.device)
        kernels.tv_loss(image.contiguous(), 1.0, 0.0, acc, None)
        return acc[0].to(image.dtype)


def _plane(mask, device):
    """(1,H,W,1) / (H,W) mask (numpy or tensor) -> contiguous float32 CUDA plane (H,W)."""
    m = torch.as_tensor(mask)
    if m.dim() == 4:
        m = m[0, :, :, 0]
    return m.to(device=device, dtype=torch.float32).contiguous()


def dict_zip(*dicts):                                                         # loss.py:163-165
    for k in dicts[0].keys():
        yield [d[k] for d in dicts]
