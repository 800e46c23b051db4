"""One large image optimised on several GPUs (BASELINE configs[3]).

The image is cut into column strips, one per rank; a rank stores its strip plus a halo of HALO = 32 image pixels on every
interior side (halo of a level-l feature map: 32 / 2^l columns: 32, 16, 8, 4, 2 for blocks 1..5).  What a rank needs from
its neighbours is exchanged, point to point over NVLink, between the SEGMENTS of the network (the five blocks, block4 cut in
two):

  forward    the tensor that leaves a segment (the pooled tensor; conv 9's output inside block4) gets its halo columns
             overwritten with the owner's values, so each segment starts from exact inputs on own + halo.  Inside a segment of
             n convolutions the valid region shrinks by one column per convolution (the local edge is zero padded); the
             backward pass needs the activations' ReLU masks / arg-max routing exact on n halo columns, so a segment may
             hold n <= halo / 2 convolutions: (2, 2, 4, 2, 2, 1) against halos of (32, 16, 8, 4, 4, 2) columns.
  backward   the gradient that flows from a segment into the one below is exact on the own columns only; its halo columns
             are overwritten with the owners' complete values before the lower segment continues.
  image      after the Adam update the HALO columns next to each interior boundary are refreshed from the owner.

Eleven exchanges per iteration (5 forward, 5 backward, 1 image, with both neighbours each), 0.1-2 MB per slab at 3840x2160; the
redundant convolution work is (own + 2 x 32) / own -- it was (own + 2 x 160) / own with round 1's overlapped strips.
The only collective is ONE NCCL all-reduce of the flattened per-class Gram partials of the five style layers (each rank sums
over its OWN pixels; 19.5 MB at K = 8) plus the 4-entry float64 loss accumulator.

Strip boundaries must be multiples of 16 px so that the pooling grids and the bilinear mask resizing of every layer align
with the global image (restricting a resized mask to the own columns is then exact).
"""
import threading

import torch

from . import kernels
from .components.VGG19.model import StyleContentModel
from .components.loss import Loss
from .style_transfer import Adam, CONTENT_LAYERS, STYLE_LAYERS

HALO = 32


class Tile:
    """Column strip of a W-pixel-wide image: own = [own_lo, own_hi), stored = [ext_lo, ext_hi) (own + halo)."""

    def __init__(self, W, rank, world, halo=HALO):
        if W % (16 * world) != 0:
            raise ValueError("image width %d must be a multiple of 16 x %d ranks" % (W, world))
        if world > 1 and W // world < halo:
            raise ValueError("strips of %d px are narrower than the %d-px halo" % (W // world, halo))
        self.W, self.rank, self.world, self.halo = W, rank, world, halo
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

    def halo_cols(self, w_layer):
        return self.halo // self._factor(w_layer)

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


# ----------------------------------------------------------------------------------------------------------------
# communication back ends: NCCL (one process per GPU) and an in-process stand-in (threads, one per emulated rank)
# ----------------------------------------------------------------------------------------------------------------
class NcclComm:
    """torch.distributed (NCCL): batched point-to-point for the halos, all-reduce for the Gram partials."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.bytes = {"allreduce": 0, "halo": 0, "exchanges": 0}

    def begin_step(self):
        self.bytes = {"allreduce": 0, "halo": 0, "exchanges": 0}

    def reduce_sum(self, tensors):
        import torch.distributed as dist
        for t in tensors:
            if self.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            self.bytes["allreduce"] += t.numel() * t.element_size()

    def exchange(self, tensor, lo, hi, hl):
        """tensor (1,h,w,C) or (h,w,C)-like with columns on dim -2: send the hl own columns next to each interior boundary to
        that neighbour, overwrite the hl halo columns with what the neighbour sends.  Returns the received slabs."""
        import torch.distributed as dist
        t = tensor if tensor.dim() == 4 else tensor.unsqueeze(0)
        ops, recv = [], []
        if self.rank > 0 and lo > 0:                                   # left neighbour
            snd = t[:, :, lo:lo + hl].contiguous()
            rcv = torch.empty_like(snd)
            ops += [dist.P2POp(dist.isend, snd, self.rank - 1), dist.P2POp(dist.irecv, rcv, self.rank - 1)]
            recv.append((rcv, lo - hl, lo))
            self.bytes["halo"] += snd.numel() * 4
        if self.rank < self.world - 1 and hi < t.shape[2]:             # right neighbour
            snd = t[:, :, hi - hl:hi].contiguous()
            rcv = torch.empty_like(snd)
            ops += [dist.P2POp(dist.isend, snd, self.rank + 1), dist.P2POp(dist.irecv, rcv, self.rank + 1)]
            recv.append((rcv, hi, hi + hl))
            self.bytes["halo"] += snd.numel() * 4
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            self.bytes["exchanges"] += 1
        for rcv, a, b in recv:
            t[:, :, a:b] = rcv
        return [r for r, _, _ in recv]


class GlooComm(NcclComm):
    """The same protocol over a gloo group with host copies: CPU tests of the exchange code (tests/test_host_logic.py)."""

    def reduce_sum(self, tensors):
        import torch.distributed as dist
        for t in tensors:
            c = t.detach().cpu()
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            t.copy_(c)
            self.bytes["allreduce"] += t.numel() * t.element_size()

    def exchange(self, tensor, lo, hi, hl):
        import torch.distributed as dist
        t = tensor if tensor.dim() == 4 else tensor.unsqueeze(0)
        recv = []
        # gloo has no batched isend/irecv on every build: order the blocking calls by parity instead
        def swap(peer, snd_cols, rcv_cols):
            snd = t[:, :, snd_cols[0]:snd_cols[1]].detach().cpu().contiguous()
            rcv = torch.empty_like(snd)
            if self.rank < peer:
                dist.send(snd, peer); dist.recv(rcv, peer)
            else:
                dist.recv(rcv, peer); dist.send(snd, peer)
            t[:, :, rcv_cols[0]:rcv_cols[1]] = rcv.to(t.device)
            recv.append(rcv.to(t.device))
            self.bytes["halo"] += snd.numel() * 4
        # even ranks talk to the right first, odd ranks to the left first (no deadlock on a chain)
        order = ("right", "left") if self.rank % 2 == 0 else ("left", "right")
        for side in order:
            if side == "left" and self.rank > 0 and lo > 0:
                swap(self.rank - 1, (lo, lo + hl), (lo - hl, lo))
            if side == "right" and self.rank < self.world - 1 and hi < t.shape[2]:
                swap(self.rank + 1, (hi - hl, hi), (hi, hi + hl))
        self.bytes["exchanges"] += 1
        return recv


class ThreadComm:
    """Several ranks inside ONE process (single-GPU emulation, tests): every rank runs in its own thread on the same CUDA
    stream; an exchange is a mailbox plus two barriers.  Host-side barriers order the enqueueing, the shared stream orders
    the device work."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.box = {}

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world
        self.bytes = {"allreduce": 0, "halo": 0, "exchanges": 0}

    def begin_step(self):
        self.bytes = {"allreduce": 0, "halo": 0, "exchanges": 0}

    def reduce_sum(self, tensors):
        self.s.box[("red", self.rank)] = tensors
        self.s.barrier.wait()
        if self.rank == 0:
            for group in zip(*[self.s.box[("red", r)] for r in range(self.world)]):
                tot = torch.stack([g.to(torch.float64) for g in group]).sum(0)
                for g in group:
                    g.copy_(tot.to(g.dtype))
        self.s.barrier.wait()
        self.bytes["allreduce"] += sum(t.numel() * t.element_size() for t in tensors)

    def exchange(self, tensor, lo, hi, hl):
        t = tensor if tensor.dim() == 4 else tensor.unsqueeze(0)
        has_left, has_right = self.rank > 0 and lo > 0, self.rank < self.world - 1 and hi < t.shape[2]
        if has_left:
            self.s.box[("to", self.rank - 1, "from_right")] = t[:, :, lo:lo + hl].clone()
        if has_right:
            self.s.box[("to", self.rank + 1, "from_left")] = t[:, :, hi - hl:hi].clone()
        self.s.barrier.wait()
        recv = []
        if has_left:
            r = self.s.box[("to", self.rank, "from_left")]
            t[:, :, lo - hl:lo] = r
            recv.append(r)
        if has_right:
            r = self.s.box[("to", self.rank, "from_right")]
            t[:, :, hi:hi + hl] = r
            recv.append(r)
        self.s.barrier.wait()
        self.bytes["halo"] += sum(r.numel() * 4 for r in recv)
        self.bytes["exchanges"] += 1
        return recv


class TiledStyleTransfer:
    """Rank-local state of a spatially tiled optimisation.  `comm` provides exchange(tensor, lo, hi, hl) and
    reduce_sum(list of tensors); the default is NCCL through torch.distributed."""

    def __init__(self, content, style, args, content_masks, style_masks, vgg_weights, rank, world, comm=None, matting="v2",
                 device=None):
        dev = torch.device(device if device is not None else "cuda")
        self.rank, self.world = rank, world
        self.comm = comm if comm is not None else NcclComm(rank, world)
        content = torch.as_tensor(content, dtype=torch.float32)
        style = torch.as_tensor(style, dtype=torch.float32)
        self.tile = Tile(int(content.shape[2]), rank, world)
        self.style_tile = Tile(int(style.shape[2]), rank, world)
        c_loc = self.tile.crop(content).to(dev)
        s_loc = self.style_tile.crop(style).to(dev)
        cm = None if content_masks is None else [self.tile.crop(torch.as_tensor(m)) for m in content_masks]
        sm = None if style_masks is None else [self.style_tile.crop(torch.as_tensor(m)) for m in style_masks]
        self.extractor = StyleContentModel(CONTENT_LAYERS, STYLE_LAYERS, shape=(None, None, 3), weights=vgg_weights, device=dev)
        # the targets are features of strips as well: same block-wise forward pass with halo exchange
        content_target = self.extractor.forward_blocks(c_loc, self._exchanger(self.tile), reuse=False)['content']
        style_target = self.extractor.forward_blocks(s_loc, self._exchanger(self.style_tile), reuse=False)['style']
        self.loss = Loss(content_target, style_target, args, cm, sm, matting=matting, tile=self.tile, style_tile=self.style_tile)
        if args.regularization_weight > 0:
            self.loss.initialize_matting_laplacian(c_loc[0].to(torch.float64))
        self.optimizer = Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon)
        self.image = c_loc.clone()                      # the local strip of the transfer image (own columns + halo)
        self._grad = torch.empty_like(self.image)
        self._targets_reduced = False
        self._flat = None                               # flat float32 buffer behind the per-layer Gram partials

    def _exchanger(self, tile):
        """exchange(tensor) for the image, feature maps and gradients of this strip (the level follows from the width)."""
        def ex(tensor):
            w_l = int(tensor.shape[-2])
            lo, hi = tile.own_cols(w_l)
            return self.comm.exchange(tensor, lo, hi, tile.halo_cols(w_l))
        return ex

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
        return ("point-to-point halo exchange between the six network segments (forward activations, backward gradients) and of "
                "the image border after the update (%d px halo), one NCCL all-reduce of the flattened Gram partials (%.1f MB) + float64[4] "
                "loss accumulator" % (HALO, 4e-6 * (self._flat.numel() if self._flat is not None else 0)))

    def exchange_bytes(self):
        """Bytes this rank handed to the communication layer in the latest step."""
        return dict(self.comm.bytes)

    def time_breakdown(self, steps=5):
        """Device time of `steps` iterations split into communication (halo exchanges incl. packing / unpacking, the Gram
        all-reduce) and everything else, measured with CUDA events around every call into `comm` (the events serialise
        nothing: communication is already stream-ordered with the compute).  Returns ms per step."""
        events = []
        comm = self.comm
        raw_exchange, raw_reduce = comm.exchange, comm.reduce_sum

        def timed(fn):
            def wrapper(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(*a, **k)
                e1.record()
                events.append((e0, e1))
                return out
            return wrapper

        comm.exchange, comm.reduce_sum = timed(raw_exchange), timed(raw_reduce)
        try:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(steps):
                self.step()
            t1.record()
            torch.cuda.synchronize()
        finally:
            comm.exchange, comm.reduce_sum = raw_exchange, raw_reduce
        total = t0.elapsed_time(t1) / steps
        comm_ms = sum(a.elapsed_time(b) for a, b in events) / steps
        own = self.tile.own_hi - self.tile.own_lo
        return {"ms_per_step": total, "communication_ms": comm_ms, "compute_ms": total - comm_ms,
                "redundant_column_factor": self.tile.local_w / own}

    def own_strip(self):
        lo, hi = self.tile.own_cols(self.tile.local_w)
        return self.image[0, :, lo:hi].contiguous()

    def step(self):
        """One iteration (every rank calls it; the calls to `comm` are collective)."""
        self.comm.begin_step()
        ex = self._exchanger(self.tile)
        outputs = self.extractor.forward_blocks(self.image, ex, reuse=True)
        if self._flat is None:
            self.loss.prepare(outputs)                  # per-layer state (idempotent), then one buffer behind all Gram partials
            self._flatten_partials()
        if not self._targets_reduced:                   # the style Grams computed at set-up are per-rank partials: sum once
            self.comm.reduce_sum(self.loss.style_targets_partial())
            self._targets_reduced = True
        parts = self.loss.forward_partials(self.image, outputs)
        self.comm.reduce_sum([self._flat, parts[-1]])   # every Gram partial lives in the flat buffer; + the float64 accumulator
        loss_dict = self.loss.finish()
        grad = self.loss.gradient(self.extractor, out=self._grad,
                                  backward=lambda seeds, out: self.extractor.backward_blocks(seeds, ex, out=out))
        self.optimizer.apply_gradients_and_clip(grad, self.image)      # only the own columns of the result are meaningful
        ex(self.image)                                                 # refresh the image halo from the owners
        return loss_dict


def run_emulated(ranks, iters):
    """Drive several TiledStyleTransfer objects that live in ONE process (their `comm` must be ThreadComm objects sharing one
    ThreadComm.Shared): one thread per rank.  Returns the list of loss dicts of rank 0 (identical on every rank)."""
    history = [[] for _ in ranks]
    errors = []

    def work(i, r):
        try:
            for _ in range(iters):
                d = r.step()
                history[i].append({k: float(v) for k, v in d.items()})
        except BaseException as e:           # make a failing rank visible instead of dead-locking the barrier
            errors.append(e)
            try:
                r.comm.s.barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(i, r)) for i, r in enumerate(ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return history[0]


def make_emulated(content, style, args, content_masks, style_masks, vgg_weights, world, matting="v2", device=None):
    """`world` rank objects in this process.  Construction itself exchanges halos (the targets), so it runs in threads too."""
    shared = ThreadComm.Shared(world)
    ranks, errors = [None] * world, []

    def build(r):
        try:
            ranks[r] = TiledStyleTransfer(content, style, args, content_masks, style_masks, vgg_weights, r, world,
                                          comm=ThreadComm(shared, r), matting=matting, device=device)
        except BaseException as e:
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=build, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return ranks
