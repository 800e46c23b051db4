"""One large image optimised on several GPUs (BASELINE configs[3]).

The image is cut into column strips, one per rank; on every resolution level l (image and block1: 0, block2: 1, ... block5: 4)
a rank stores its own columns plus LEVEL_HALO[l] = (4, 4, 8, 4, 2) halo columns on every interior side.  What a rank needs
from its neighbours is exchanged, point to point over NVLink, between the SEGMENTS of the network (the five blocks, block4
cut in two):

  forward    the tensor that leaves a segment (the pooled tensor; conv 9's output inside block4) gets its halo columns
             overwritten with the owner's values, so each segment starts from exact inputs on own + halo.  Inside a segment of
             n convolutions the valid region shrinks by one column per convolution (the local edge is zero padded); the
             backward pass needs the activations' ReLU masks / arg-max routing exact on n halo columns, so a segment may
             hold n <= halo / 2 convolutions: (2, 2, 4, 2, 2, 1) against halos of (4, 4, 8, 4, 4, 2) columns.
  backward   the gradient that flows from a segment into the one below is exact on the own columns only; its halo columns
             are overwritten with the owners' complete values before the lower segment continues.
  image      after the Adam update the halo columns of the image are refreshed from the owner.

Because the halo of a pooled tensor is overwritten by the neighbour anyway, the halos of two levels are independent: the pool
of a level-l tensor (own/2 + h_l/2 columns per side) is written INTO the wider level-(l+1) tensor at a column offset
(adpst_vgg_forward_range's level_widths / pool_col_offset), and levels 0 and 1 -- where most of the work is -- carry 4 columns
instead of the 32 / 16 that a uniform 2^-l geometry needs for block5's two.  Redundant work at 8 GPUs and 3840 px: 1.7 % on
block1, 3.3 % on block2, 13 % on blocks 3-5 (round 1's overlapped strips: 67 % everywhere).

Eleven exchanges per iteration (5 forward, 5 backward, 1 image, with both neighbours each), 0.1-2 MB per slab at 3840x2160.
The only collective is the NCCL all-reduce of the per-class Gram partials of the five style layers (each rank sums over its
OWN pixels; 19.5 MB at K = 8) plus the 4-entry float64 loss accumulator.

Strip boundaries must be multiples of 16 px so that the pooling grids and the bilinear mask resizing of every layer align
with the global image (restricting a resized mask to the own columns is then exact).
"""
import ctypes
import threading
import time

import torch

from . import _lib, kernels
from .components.VGG19.model import POOL_AFTER, SEGMENTS, StyleContentModel, VGG19Handle
from .components.loss import Loss
from .style_transfer import Adam, CONTENT_LAYERS, STYLE_LAYERS

LEVEL_HALO = (4, 4, 8, 4, 2)            # halo columns per interior side on resolution level 0..4
HALO = LEVEL_HALO[0]                    # ... of the image


def _level_convs():
    """For every resolution level: the sizes of the network segments on it (number of convolutions between two exchanges)."""
    out = [[] for _ in range(5)]
    for first, last in SEGMENTS:
        out[sum(1 for p in POOL_AFTER if p < first)].append(last - first + 1)
    return out


class Tile:
    """Column strip of a W-pixel-wide image on every resolution level: own columns [own_lo, own_hi) of the image, and per level
    l the local tensor layout  [left halo | own >> l | right halo]  with halos[l] columns on interior sides, none at the
    image border."""

    def __init__(self, W, rank, world, halos=LEVEL_HALO):
        if W % (16 * world) != 0:
            raise ValueError("image width %d must be a multiple of 16 x %d ranks" % (W, world))
        halos = tuple(int(h) for h in halos)
        self.W, self.rank, self.world, self.halos = W, rank, world, halos
        self.own_lo, self.own_hi = rank * W // world, (rank + 1) * W // world
        own = self.own_hi - self.own_lo
        convs = _level_convs()
        for l in range(5):
            if world > 1 and (own >> l) < halos[l]:
                raise ValueError("strips of %d px are too narrow for a %d-column halo on level %d" % (own, halos[l], l))
            if any(halos[l] < 2 * n for n in convs[l]):
                raise ValueError("level %d: %d halo columns cannot carry segments of %s convolutions" % (l, halos[l], convs[l]))
            if l < 4 and (halos[l] % 2 or 2 * halos[l + 1] < convs[l][-1]):
                raise ValueError("level %d: halos %s do not line up with the pooling grid" % (l, halos))
        self.has_left, self.has_right = rank > 0, rank < world - 1
        self.left = tuple(h if self.has_left else 0 for h in halos)
        self.right = tuple(h if self.has_right else 0 for h in halos)
        self.widths = tuple(self.left[l] + (own >> l) + self.right[l] for l in range(5))
        if len(set(self.widths)) != 5:
            raise ValueError("strip geometry %s: two levels have the same width" % (self.widths,))
        # column where the 2x2 pool of a level-l tensor starts inside the level-(l+1) tensor
        self.pool_xoff = tuple(self.left[l + 1] - self.left[l] // 2 for l in range(4))
        self.halo = halos[0]
        self.ext_lo, self.ext_hi = self.own_lo - self.left[0], self.own_hi + self.right[0]
        self.local_w = self.widths[0]
        # the segmentation masks are kept at full resolution for the widest footprint of any level
        reach = max(h << l for l, h in enumerate(halos))
        self.mask_lo = self.own_lo - (reach if self.has_left else 0)
        self.mask_hi = self.own_hi + (reach if self.has_right else 0)

    def geom(self):
        """(level widths, pool column offsets) for adpst_vgg_forward_range / _backward_range; None for a single strip."""
        return None if self.world == 1 else (self.widths, self.pool_xoff)

    def level(self, w_layer):
        try:
            return self.widths.index(int(w_layer))
        except ValueError:
            raise ValueError("no level of this strip is %d columns wide (%s)" % (w_layer, self.widths))

    def own_cols(self, w_layer):
        """Own column range in the coordinates of a tensor that is w_layer columns wide."""
        l = self.level(w_layer)
        return self.left[l], self.left[l] + ((self.own_hi - self.own_lo) >> l)

    def global_cols(self, w_layer):
        return self.W >> self.level(w_layer)

    def halo_cols(self, w_layer):
        return self.halos[self.level(w_layer)]

    def mask_cols(self, w_layer):
        """Columns of the cropped full-resolution masks (crop_masks) that a tensor of w_layer columns covers."""
        l = self.level(w_layer)
        a = (self.own_lo - (self.left[l] << l)) - self.mask_lo
        return a, a + (self.widths[l] << l)

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

    def crop_masks(self, mask):
        """mask (H, W) or (1, H, W, 1) (mask_for_tf): the columns kept at full resolution, same rank."""
        if mask.dim() == 4:
            return mask[:, :, self.mask_lo:self.mask_hi].contiguous()
        return mask[:, self.mask_lo:self.mask_hi].contiguous()


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

    def reduce_start(self, t):
        """Start summing t over the ranks on NCCL's own stream (it waits for the work queued on the current stream so far) and
        return a callable that makes the current stream wait for the result.  Kernels enqueued in between overlap with it."""
        import torch.distributed as dist
        self.bytes["allreduce"] += t.numel() * t.element_size()
        if self.world == 1:
            return lambda: None
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True).wait

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

    def reduce_start(self, t):
        self.reduce_sum([t])
        return lambda: None

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

    def reduce_start(self, t):
        self.reduce_sum([t])
        return lambda: None

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


def exchange_plan(H, tile, last_layer):
    """(rows, halo columns, channels) of the tensors whose halos are exchanged in one step, in order: the tensors that leave
    the network segments on the way up, the gradients w.r.t. the same tensors on the way down, the image."""
    up = []
    for first, last in SEGMENTS:
        if last >= last_layer:
            break
        level = sum(1 for p in POOL_AFTER if p <= last)
        c = VGG19Handle.conv_shape(last, H, tile.local_w)[2]
        up.append((H >> level, tile.halos[level], c))
    return up + up[::-1] + [(H, tile.halos[0], 3)]


class PeerHalo:
    """Halo exchange through mailboxes in peer memory (csrc/halo_peer.cu): the sender's kernel stores its boundary columns
    straight into the neighbour's mailbox over NVLink and releases a sequence number, the receiver's kernel waits for it and
    copies the slab into its halo columns (folding max|slab| into the tensor's scale word).  Two launches per exchange, no
    host involvement, capturable in a CUDA graph."""

    def __init__(self, side_bytes, rank, world, device=None):
        _lib.require_cuda()
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.side_bytes = (int(side_bytes) + 15) // 16 * 16
        self.sync = None                     # ranks of one process on one stream: called between the pushes and the pulls
        self._slot, self._off = 0, 0
        self.bytes = {"halo": 0, "exchanges": 0}
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().adpst_halo_create(self.side_bytes, ctypes.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().adpst_halo_destroy(h)
            except Exception:
                pass

    def connect_processes(self):
        """Collective over the default process group (one process per GPU of one node): map the neighbours' mailboxes
        through CUDA IPC.  Every rank raises RuntimeError if any rank could not connect (the ranks agree on the outcome, so
        the caller can fall back to another transport collectively)."""
        import torch.distributed as dist
        L = _lib.lib()
        buf = ctypes.create_string_buffer(L.adpst_halo_ipc_handle_bytes())
        error = None
        try:
            _lib.check(L.adpst_halo_export(self._h, buf))
        except _lib.AdpstError as e:
            error = str(e)
        mine = (self.rank, bytes(buf.raw), self.side_bytes, error)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        table = {r: (hb, sb) for r, hb, sb, _ in everyone}
        error = next((e for _, _, _, e in everyone if e), None)
        if error is None:
            try:
                with torch.cuda.device(self.device):
                    for side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
                        if 0 <= peer < self.world:
                            hb, sb = table[peer]
                            _lib.check(L.adpst_halo_connect_ipc(self._h, side, ctypes.c_char_p(hb), sb))
            except _lib.AdpstError as e:
                error = str(e)
        outcomes = [None] * self.world
        dist.all_gather_object(outcomes, error)
        failed = [(r, e) for r, e in enumerate(outcomes) if e]
        if failed:
            raise RuntimeError("peer-memory mailboxes could not be connected (rank %d: %s)" % failed[0])

    @staticmethod
    def connect_local(halos, sync=None):
        """Ranks that live in one process: neighbours use each other's mailbox pointers directly."""
        L = _lib.lib()
        for r, h in enumerate(halos):
            with torch.cuda.device(h.device):
                if r > 0:
                    _lib.check(L.adpst_halo_connect_local(h._h, 0, halos[r - 1]._h))
                if r + 1 < len(halos):
                    _lib.check(L.adpst_halo_connect_local(h._h, 1, halos[r + 1]._h))
            h.sync = sync

    def begin_step(self):
        self._slot, self._off = 0, 0
        self.bytes = {"halo": 0, "exchanges": 0}

    def exchange(self, tensor, lo, hi, hl, slot_ptr=None, overlap=None):
        """tensor (1, rows, width, C) float32, own columns [lo, hi): see the class comment.  slot_ptr: device address of the
        tensor's scale word or None.  overlap: callable that enqueues independent work between the push and the pull, so that
        the slabs travel (and a slower neighbour catches up) while this GPU keeps computing."""
        if tensor.dim() != 4 or tensor.shape[0] != 1 or tensor.dtype != torch.float32 or not tensor.is_contiguous():
            raise TypeError("expected a contiguous float32 tensor of shape (1, rows, width, C)")
        rows, width, C = int(tensor.shape[1]), int(tensor.shape[2]), int(tensor.shape[3])
        L = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(L.adpst_halo_push(self._h, self._slot, self._off, _lib.ptr(tensor), rows, width, C, hl, lo, hi,
                                         _lib.stream_ptr()))
            if self.sync is not None:
                self.sync()
            if overlap is not None:
                overlap()
            _lib.check(L.adpst_halo_pull(self._h, self._slot, self._off, _lib.ptr(tensor), rows, width, C, hl, lo, hi,
                                         ctypes.c_void_p(slot_ptr or 0), _lib.stream_ptr()))
            if self.sync is not None:
                self.sync()
        sides = int(self.rank > 0 and lo > 0) + int(self.rank < self.world - 1 and hi < width)
        self.bytes["halo"] += sides * rows * hl * C * 4
        self.bytes["exchanges"] += 1
        self._slot += 1
        self._off += rows * hl * C * 4


class TiledStyleTransfer:
    """Rank-local state of a spatially tiled optimisation.  `comm` provides exchange(tensor, lo, hi, hl) and
    reduce_sum(list of tensors); the default is NCCL through torch.distributed.  halo: None -- the halos travel through
    `comm` as well; 'peer' -- through a PeerHalo mailbox connected over CUDA IPC (collective); or a connected PeerHalo.  The
    set-up passes (content / style targets) always use `comm`."""

    def __init__(self, content, style, args, content_masks, style_masks, vgg_weights, rank, world, comm=None, matting="v2",
                 device=None, halo=None, overlap_reduce=True):
        dev = torch.device(device if device is not None else "cuda")
        self.rank, self.world = rank, world
        self.comm = comm if comm is not None else NcclComm(rank, world)
        content = torch.as_tensor(content, dtype=torch.float32)
        style = torch.as_tensor(style, dtype=torch.float32)
        self.tile = Tile(int(content.shape[2]), rank, world)
        self.style_tile = Tile(int(style.shape[2]), rank, world)
        c_loc = self.tile.crop(content).to(dev)
        s_loc = self.style_tile.crop(style).to(dev)
        cm = None if content_masks is None else [self.tile.crop_masks(torch.as_tensor(m)) for m in content_masks]
        sm = None if style_masks is None else [self.style_tile.crop_masks(torch.as_tensor(m)) for m in style_masks]
        self.extractor = StyleContentModel(CONTENT_LAYERS, STYLE_LAYERS, shape=(None, None, 3), weights=vgg_weights, device=dev)
        # the targets are features of strips as well: same block-wise forward pass with halo exchange
        content_target = self.extractor.forward_blocks(c_loc, self._exchanger(self.tile), reuse=False,
                                                       geom=self.tile.geom())['content']
        style_target = self.extractor.forward_blocks(s_loc, self._exchanger(self.style_tile), reuse=False,
                                                     geom=self.style_tile.geom())['style']
        self.loss = Loss(content_target, style_target, args, cm, sm, matting=matting, tile=self.tile, style_tile=self.style_tile)
        if args.regularization_weight > 0:
            self.loss.initialize_matting_laplacian(c_loc[0].to(torch.float64))
        self.optimizer = Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon)
        self.image = c_loc.clone()                      # the local strip of the transfer image (own columns + halo)
        self._grad = torch.empty_like(self.image)
        self._targets_reduced = False
        self._flat = None                               # flat float32 buffer behind the per-layer Gram partials
        li = self.extractor.last_index
        self._last_exchange_layer = max(last for _, last in SEGMENTS if last < li)     # the forward pass's last exchange
        self.overlap_reduce = overlap_reduce            # sum finished Gram partials while the forward pass continues
        self.halo = None
        if halo is not None and world > 1:
            if isinstance(halo, str):
                if halo != "peer":
                    raise ValueError("halo must be None, 'peer' or a PeerHalo")
                halo = PeerHalo(self.mailbox_bytes(), rank, world, dev)
                halo.connect_processes()
            if halo.side_bytes < self.mailbox_bytes():
                raise ValueError("the mailbox holds %d bytes per side, one step needs %d" % (halo.side_bytes, self.mailbox_bytes()))
            self.halo = halo

    def mailbox_bytes(self):
        """Bytes per mailbox side that the exchanges of one step occupy."""
        plan = exchange_plan(int(self.image.shape[1]), self.tile, self.extractor.last_index)
        return sum(r * h * c * 4 for r, h, c in plan)

    def _exchanger(self, tile, peer=False):
        """exchange(tensor, slot) for the image, feature maps and gradients of this strip (the level follows from the width):
        halo columns replaced by the neighbours' own columns, the scale word at device address `slot` raised accordingly."""
        vgg = self.extractor.vgg

        def ex(tensor, slot=None, overlap=None):
            w_l = int(tensor.shape[-2])
            lo, hi = tile.own_cols(w_l)
            if peer and self.halo is not None:
                self.halo.exchange(tensor, lo, hi, tile.halo_cols(w_l), slot, overlap)
                return
            if overlap is not None:
                overlap()
            for slab in self.comm.exchange(tensor, lo, hi, tile.halo_cols(w_l)):
                if slot:
                    vgg.absmax_update(slab, slot)
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
        how = ("two kernels per exchange over peer memory: stores into the neighbour's mailbox through NVLink + release flag, "
               "then wait + copy into the halo columns" if self.halo is not None else "NCCL batched send/recv")
        return ("halo exchange between the six network segments (forward activations, backward gradients) and of the image "
                "border after the update (halo columns per resolution level %s; %s), NCCL all-reduce of the Gram partials (%.1f MB, "
                "started as each layer's partial is complete) + float64[4] loss accumulator"
                % (list(self.tile.halos), how, 4e-6 * (self._flat.numel() if self._flat is not None else 0)))

    def exchange_bytes(self):
        """Bytes this rank handed to the communication layer in the latest step."""
        b = dict(self.comm.bytes)
        if self.halo is not None:
            b["halo"] += self.halo.bytes["halo"]
            b["exchanges"] += self.halo.bytes["exchanges"]
        return b

    def time_breakdown(self, steps=5):
        """Device time of `steps` iterations split into communication (halo exchanges incl. packing / unpacking, the Gram
        all-reduce) and everything else, measured with CUDA events around every call into `comm` (the events serialise
        nothing: communication is already stream-ordered with the compute).  Returns ms per step."""
        events = []
        comm = self.comm
        raw_exchange, raw_reduce = comm.exchange, comm.reduce_sum
        halo = self.halo
        raw_peer = None if halo is None else halo.exchange

        reduces = []

        def timed(fn, into=None):
            into = events if into is None else into

            def wrapper(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(*a, **k)
                e1.record()
                into.append((e0, e1))
                return out
            return wrapper

        hidden = []                                  # (start, end) of the work enqueued inside an exchange window

        def timed_peer(tensor, lo, hi, hl, slot_ptr=None, overlap=None):
            def inner():
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                overlap()
                b.record()
                hidden.append((a, b))
            return raw_peer(tensor, lo, hi, hl, slot_ptr, None if overlap is None else inner)

        comm.exchange, comm.reduce_sum = timed(raw_exchange), timed(raw_reduce, reduces)
        if halo is not None:
            halo.exchange = timed(timed_peer)
        try:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            h0 = time.perf_counter()
            t0.record()
            for _ in range(steps):
                self.step()
            t1.record()
            host_ms = (time.perf_counter() - h0) * 1e3 / steps
            torch.cuda.synchronize()
        finally:
            comm.exchange, comm.reduce_sum = raw_exchange, raw_reduce
            if halo is not None:
                del halo.exchange                   # back to the class's method
        total = t0.elapsed_time(t1) / steps
        reduce_ms = sum(a.elapsed_time(b) for a, b in reduces) / steps
        comm_ms = (sum(a.elapsed_time(b) for a, b in events) - sum(a.elapsed_time(b) for a, b in hidden)) / steps + reduce_ms
        own = self.tile.own_hi - self.tile.own_lo
        return {"ms_per_step": total, "communication_ms": comm_ms, "allreduce_ms": reduce_ms, "compute_ms": total - comm_ms,
                "host_enqueue_ms": host_ms, "redundant_column_factor": self.tile.local_w / own,
                "redundant_column_factor_blocks345": self.tile.widths[2] / (own >> 2)}

    def own_strip(self):
        lo, hi = self.tile.own_cols(self.tile.local_w)
        return self.image[0, :, lo:hi].contiguous()

    def graphed_step(self):
        """Capture one iteration -- kernels, the NCCL point-to-point exchanges and the all-reduce -- into a CUDA graph and
        return replay() -> loss dict.  Collective: every rank captures, and every rank replays the same number of times.
        At least one eager step() must have run (lazy state, NCCL communicators, persistent buffers).  The capture runs in
        thread-local error mode: NCCL's watchdog thread may query its events while this thread captures."""
        if self._flat is None or not self._targets_reduced:
            raise RuntimeError("graphed_step() needs a preceding eager step()")
        if not isinstance(self.comm, NcclComm) or isinstance(self.comm, GlooComm):
            raise RuntimeError("graphed_step() needs the NCCL communication layer")
        s = _capture_stream(self.image.device)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            g.capture_begin(capture_error_mode="thread_local")
            try:
                out = self.step()
            finally:
                g.capture_end()
        torch.cuda.current_stream().wait_stream(s)
        self._graph = g                                 # keeps the graph's private pool (send / receive slabs) alive

        def replay():
            g.replay()
            return out
        return replay

    def step(self):
        """One iteration (every rank calls it; the calls to `comm` are collective)."""
        self.comm.begin_step()
        if self.halo is not None:
            self.halo.begin_step()
        ex = self._exchanger(self.tile, peer=True)
        loss = self.loss
        first_step = self._flat is None
        # Loss terms are enqueued as soon as their inputs exist, inside the exchange windows of the forward pass (Gram
        # partials and the content term of the layers below the exchanged tensor; the image-only terms in the last window).
        # On the first step the per-layer state does not exist yet: everything runs after the forward pass.
        loss.begin_partials(self.image)

        # A Gram partial that is complete is summed over the ranks at once, on NCCL's stream, while this rank computes the
        # deeper segments; only the last layer's partial and the float64 accumulator are reduced after the forward pass.
        waits, reduced = [], set()

        def fwd_overlap(last, outputs):
            loss.partial_layers(outputs, last)
            if self.overlap_reduce:
                for name, G in loss._part["G"].items():
                    if name not in reduced:
                        reduced.add(name)
                        waits.append(self.comm.reduce_start(G))
            if last >= self._last_exchange_layer:
                loss.partial_image()
        outputs = self.extractor.forward_blocks(self.image, ex, reuse=True, overlap=None if first_step else fwd_overlap,
                                                geom=self.tile.geom())
        if first_step:
            loss.prepare(outputs)                       # per-layer state (idempotent), then one buffer behind all Gram partials
            self._flatten_partials()
        if not self._targets_reduced:                   # the style Grams computed at set-up are per-rank partials: sum once
            self.comm.reduce_sum(loss.style_targets_partial())
            self._targets_reduced = True
        parts = loss.end_partials(outputs)
        if reduced:                                     # the partials still local (views of the flat buffer) + the accumulator
            self.comm.reduce_sum([G for name, G in loss._part["G"].items() if name not in reduced] + [parts[-1]])
            for wait in waits:
                wait()
        else:
            self.comm.reduce_sum([self._flat, parts[-1]])   # every Gram partial lives in the flat buffer
        # Style terms / gradient seeds are evaluated lazily on the way down: the seeds of a segment at the latest right before
        # it runs, one more layer (the deepest pending one) inside every exchange window.
        loss.finish(lazy=True)

        def bwd_overlap(first):
            if first >= 0:
                loss.seed_layers(first)
            else:
                loss.seed_layers(0, at_most=1)
        grad = loss.gradient(self.extractor, out=self._grad,
                             backward=lambda seeds, out: self.extractor.backward_blocks(seeds, ex, out=out, overlap=bwd_overlap))
        loss_dict = loss.finalize()
        self.optimizer.apply_gradients_and_clip(grad, self.image)      # only the own columns of the result are meaningful
        ex(self.image)                                                 # refresh the image halo from the owners
        return loss_dict


def _capture_stream(device):
    from .style_transfer import _side_stream
    return _side_stream(device)


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


def make_emulated(content, style, args, content_masks, style_masks, vgg_weights, world, matting="v2", device=None, halo=None):
    """`world` rank objects in this process.  Construction itself exchanges halos (the targets), so it runs in threads too.
    halo='peer': the steps exchange their halos through PeerHalo mailboxes (the kernels of csrc/halo_peer.cu, with the
    neighbours' mailboxes addressed directly) instead of the in-process stand-in."""
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
    if halo == "peer" and world > 1:
        need = max(r.mailbox_bytes() for r in ranks)
        boxes = [PeerHalo(need, r, world, ranks[r].image.device) for r in range(world)]
        PeerHalo.connect_local(boxes, sync=shared.barrier.wait)
        for r, b in zip(ranks, boxes):
            r.halo = b
    elif halo is not None and world > 1:
        raise ValueError("halo must be None or 'peer'")
    return ranks
