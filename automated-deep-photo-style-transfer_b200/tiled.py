"""One large image optimised on several GPUs (BASELINE configs[3]).

The image is cut into column strips, one per rank.  Each rank works on its strip extended by a halo of HALO = 160
pixels on every interior side: twice the receptive-field radius (78 px) of block5_conv1, rounded up to the 16-pixel
pooling alignment.  With that halo
  * every VGG feature that influences the gradient of an OWN pixel is computed from true image data, so no per-layer
    halo exchange is needed (redundant convolution work in the halo instead of 25 exchanges per iteration);
  * the matting Laplacian rows of the own pixels (5x5 footprint) are exact as well.
What does cross the NVLink fabric, per iteration:
  * the per-class Gram partials of the five style layers (each rank sums over its OWN pixels only): ONE NCCL all-reduce of
    one flat float32 buffer (the per-layer Gram tensors are views of it; 19.5 MB at K = 8), plus the 4-entry float64 loss
    accumulator;
  * the updated border columns: every rank sends the HALO columns next to each of its interior boundaries to that neighbour
    and receives the neighbour's (NCCL point-to-point, batched; H x 160 x 3 floats per side) -- not whole strips.
Strip boundaries must be multiples of 16 px so that pooling grids and the bilinear mask resizing of every layer align
with the global image (then restricting a resized mask to the own columns is exact).
"""
import torch

from . import kernels
from .components.VGG19.model import StyleContentModel
from .components.loss import Loss
from .style_transfer import Adam, CONTENT_LAYERS, STYLE_LAYERS

HALO = 160


class Tile:
    """Column strip of a W-pixel-wide image: own = [own_lo, own_hi), stored = [ext_lo, ext_hi) (own + halo)."""

    def __init__(self, W, rank, world, halo=HALO):
        if W % (16 * world) != 0:
            raise ValueError("image width %d must be a multiple of 16 x %d ranks" % (W, world))
        self.W, self.rank, self.world = W, rank, world
        self.own_lo, self.own_hi = rank * W // world, (rank + 1) * W // world
        self.ext_lo, self.ext_hi = max(0, self.own_lo - halo), min(W, self.own_hi + halo)
        self.local_w = self.ext_hi - self.ext_lo

    def _factor(self, w_layer):
        f = self.local_w // w_layer
        if f * w_layer != self.local_w or f & (f - 1):
            raise ValueError("layer width %d does not divide the strip width %d by a power of two" % (w_layer, self.local_w))
        return f

    def own_cols(self, w_layer):
        """Own column range in the coordinates of a layer whose local width is w_layer."""
        f = self._factor(w_layer)
        return (self.own_lo - self.ext_lo) // f, (self.own_hi - self.ext_lo) // f

    def global_cols(self, w_layer):
        return self.W // self._factor(w_layer)

    def own_masks(self, masks, K, h, w, device):
        """masks (K, h*w) or None  ->  masks restricted to the own columns (K, h*w)."""
        lo, hi = self.own_cols(w)
        ind = torch.zeros(h, w, dtype=torch.float32, device=device)
        ind[:, lo:hi] = 1.0
        if masks is None:
            return ind.reshape(1, h * w).contiguous()
        return (masks.reshape(K, h, w) * ind).reshape(K, h * w).contiguous()

    def crop(self, full_nhwc):
        return full_nhwc[:, :, self.ext_lo:self.ext_hi].contiguous()


class TiledStyleTransfer:
    """Rank-local state of a spatially tiled optimisation.  `reduce_sum(list_of_tensors)` sums each tensor over all ranks
    in place; `gather(strip)` returns the list of every rank's own strip.  Defaults use torch.distributed (NCCL)."""

    def __init__(self, content, style, args, content_masks, style_masks, vgg_weights, rank, world, reduce_sum=None,
                 gather=None, matting="v2", device=None):
        dev = torch.device(device if device is not None else "cuda")
        self.rank, self.world = rank, world
        content = torch.as_tensor(content, dtype=torch.float32)
        style = torch.as_tensor(style, dtype=torch.float32)
        self.tile = Tile(int(content.shape[2]), rank, world)
        self.style_tile = Tile(int(style.shape[2]), rank, world)
        self.reduce_sum = reduce_sum or _nccl_reduce_sum
        self.gather = gather                            # None: point-to-point border exchange over NCCL (the default)
        c_loc = self.tile.crop(content).to(dev)
        s_loc = self.style_tile.crop(style).to(dev)
        cm = None if content_masks is None else [self.tile.crop(torch.as_tensor(m)) for m in content_masks]
        sm = None if style_masks is None else [self.style_tile.crop(torch.as_tensor(m)) for m in style_masks]
        self.extractor = StyleContentModel(CONTENT_LAYERS, STYLE_LAYERS, shape=(None, None, 3), weights=vgg_weights, device=dev)
        content_target = self.extractor(c_loc)['content']
        style_target = self.extractor(s_loc)['style']
        self.loss = Loss(content_target, style_target, args, cm, sm, matting=matting, tile=self.tile, style_tile=self.style_tile)
        if args.regularization_weight > 0:
            self.loss.initialize_matting_laplacian(c_loc[0].to(torch.float64))
        self.optimizer = Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon)
        self.image = c_loc.clone()                      # the local strip of the transfer image (own columns + halo)
        self._grad = torch.empty_like(self.image)
        self._targets_reduced = False
        self._flat = None                               # flat float32 buffer behind the per-layer Gram partials
        self._bytes = {"allreduce": 0, "halo": 0}

    def _flatten_partials(self):
        """Make the per-layer transfer Grams views of ONE buffer, so that a single all-reduce sums them all."""
        sts = list(self.loss._layer_cache.values())
        n = sum(st["G"].numel() for st in sts)
        self._flat = torch.zeros(n, dtype=torch.float32, device=self.image.device)
        o = 0
        for st in sts:
            g = st["G"]
            st["G"] = self._flat[o:o + g.numel()].view(g.shape)
            o += g.numel()

    def describe_exchange(self):
        if self.world == 1:
            return "no exchange (single strip)"
        return ("one NCCL all-reduce of the flattened Gram partials (%.1f MB) + float64[4] loss accumulator, point-to-point "
                "exchange of the %d-px border columns with both neighbours" % (4e-6 * (self._flat.numel() if self._flat is not None else 0), HALO))

    def exchange_bytes(self):
        """Bytes this rank hands to NCCL per step: all-reduce payload and halo columns sent."""
        return dict(self._bytes)

    # -- the three phases of one iteration; a multi-rank driver interleaves the reductions between them ------------
    def phase_partials(self):
        outputs = self.extractor(self.image, reuse=True)
        if self._flat is None:
            self.loss.prepare(outputs)                  # per-layer state (idempotent), then one buffer behind all Gram partials
            self._flatten_partials()
        parts = self.loss.forward_partials(self.image, outputs)
        return [self._flat, parts[-1]]                  # every Gram partial lives in the flat buffer; + the float64 accumulator

    def phase_finish(self):
        loss_dict = self.loss.finish()
        grad = self.loss.gradient(self.extractor, out=self._grad)
        self.optimizer.apply_gradients_and_clip(grad, self.image)      # only the own columns of the result are meaningful
        return loss_dict

    def own_strip(self):
        lo, hi = self.tile.own_cols(self.tile.local_w)
        return self.image[0, :, lo:hi].contiguous()

    def refresh_halo(self, strips):
        """strips[r]: own strip (H, W/world, 3) of rank r after the update."""
        t = self.tile
        for r, s in enumerate(strips):
            lo, hi = r * t.W // t.world, (r + 1) * t.W // t.world
            a, b = max(lo, t.ext_lo), min(hi, t.ext_hi)
            if r != self.rank and a < b:
                self.image[0, :, a - t.ext_lo:b - t.ext_lo] = s[:, a - lo:b - lo]

    def exchange_borders(self):
        """Halo refresh over NCCL point-to-point: send the HALO own columns next to each interior boundary to that neighbour,
        receive the neighbour's into the halo columns.  Strips narrower than the halo would need data from farther ranks;
        that case falls back to gathering whole strips."""
        import torch.distributed as dist
        t = self.tile
        own_w = t.own_hi - t.own_lo
        if own_w < HALO:
            out = [torch.empty_like(self.own_strip()) for _ in range(self.world)]
            dist.all_gather(out, self.own_strip())
            self._bytes["halo"] = out[0].numel() * 4
            self.refresh_halo(out)
            return
        lo, hi = t.own_lo - t.ext_lo, t.own_hi - t.ext_lo               # own columns in local coordinates
        ops, recv = [], []
        sent = 0
        if self.rank > 0:                                               # left neighbour
            snd = self.image[0, :, lo:lo + HALO].contiguous()
            rcv = torch.empty_like(self.image[0, :, 0:lo])
            ops += [dist.P2POp(dist.isend, snd, self.rank - 1), dist.P2POp(dist.irecv, rcv, self.rank - 1)]
            recv.append((rcv, 0, lo)); sent += snd.numel() * 4
        if self.rank < self.world - 1:                                  # right neighbour
            snd = self.image[0, :, hi - HALO:hi].contiguous()
            rcv = torch.empty_like(self.image[0, :, hi:t.local_w])
            ops += [dist.P2POp(dist.isend, snd, self.rank + 1), dist.P2POp(dist.irecv, rcv, self.rank + 1)]
            recv.append((rcv, hi, t.local_w)); sent += snd.numel() * 4
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for rcv, a, b in recv:
            self.image[0, :, a:b] = rcv
        self._bytes["halo"] = sent

    def step(self):
        """One iteration with real collectives (one process per GPU)."""
        if not self._targets_reduced:
            outputs = self.extractor(self.image, reuse=True)
            self.loss.prepare(outputs)
            self.reduce_sum(self.loss.style_targets_partial())
            self._targets_reduced = True
        parts = self.phase_partials()
        self._bytes["allreduce"] = sum(p.numel() * p.element_size() for p in parts)
        self.reduce_sum(parts)
        loss_dict = self.phase_finish()
        if self.gather is not None:
            self.refresh_halo(self.gather(self.own_strip()))
        elif self.world > 1:
            self.exchange_borders()
        return loss_dict


def _nccl_reduce_sum(tensors):
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def _nccl_gather(strip):
    import torch.distributed as dist
    out = [torch.empty_like(strip) for _ in range(dist.get_world_size())]
    dist.all_gather(out, strip)
    return out


def _gloo_reduce_sum(tensors):
    """Host-side stand-in used by the CPU protocol tests: the same call sequence over a gloo group."""
    import torch.distributed as dist
    for t in tensors:
        c = t.detach().cpu()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        t.copy_(c)


def run_emulated(ranks, iters):
    """Drive several TiledStyleTransfer objects that live in ONE process (tests, single-GPU emulation): the reductions are
    plain sums over the objects.  Returns the list of loss dicts (identical on every rank)."""
    def reduce_lists(lists):
        for group in zip(*lists):
            tot = torch.stack([g.to(torch.float64) for g in group]).sum(0)
            for g in group:
                g.copy_(tot.to(g.dtype))
    history = []
    if not ranks[0]._targets_reduced:
        for r in ranks:
            r.loss.prepare(r.extractor(r.image, reuse=True))
        reduce_lists([r.loss.style_targets_partial() for r in ranks])
        for r in ranks:
            r._targets_reduced = True
    for _ in range(iters):
        reduce_lists([r.phase_partials() for r in ranks])
        dicts = [r.phase_finish() for r in ranks]
        strips = [r.own_strip() for r in ranks]
        for r in ranks:
            r.refresh_halo(strips)
        history.append({k: float(v) for k, v in dicts[0].items()})
    return history
