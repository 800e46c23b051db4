"""Oracle: VGG19 extractor, losses, Adam step.  torch-CPU, float64 by default.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/components/VGG19/model.py:4-41, components/loss.py:53-137,157-165 and
style_transfer.py:321-343.  The arithmetic of Keras VGG19 / tf.image.resize / tf.optimizers.Adam is
third-party (tensorflow, unpinned in requirements.txt:4) and is restated from its documented
behaviour: "parity unpinned".
"""
import numpy as np
import torch
import torch.nn.functional as F

# Keras applications.vgg19 topology up to block5_conv1: (name, Cin, Cout); 'P' = 2x2/2 max-pool VALID.
VGG_TOPOLOGY = [
    ("block1_conv1", 3, 64), ("block1_conv2", 64, 64), "P",
    ("block2_conv1", 64, 128), ("block2_conv2", 128, 128), "P",
    ("block3_conv1", 128, 256), ("block3_conv2", 256, 256), ("block3_conv3", 256, 256),
    ("block3_conv4", 256, 256), "P",
    ("block4_conv1", 256, 512), ("block4_conv2", 512, 512), ("block4_conv3", 512, 512),
    ("block4_conv4", 512, 512), "P",
    ("block5_conv1", 512, 512),
]
CAFFE_MEAN_BGR = (103.939, 116.779, 123.68)
CONTENT_LAYERS = ["block4_conv2"]                                  # style_transfer.py:295
STYLE_LAYERS = ["block%d_conv1" % (i + 1) for i in range(5)]       # style_transfer.py:296


def vgg_forward(image_nhwc, weights, dtype=torch.float64):
    """VGG19/model.py:27-41.  image (1,H,W,3) RGB in [0,1] -> dict name -> (1,h,w,C) post-ReLU.
    weights: dict name -> (kernel HWIO (3,3,Cin,Cout), bias (Cout,))."""
    x = image_nhwc.to(dtype) * 255.0                                # :28
    x = x.flip(-1) - torch.tensor(CAFFE_MEAN_BGR, dtype=dtype)      # :29 caffe mode: RGB->BGR, -mean
    x = x.permute(0, 3, 1, 2)
    out = {}
    for item in VGG_TOPOLOGY:
        if item == "P":
            x = F.max_pool2d(x, 2, 2)
            continue
        name, _, _ = item
        k, b = weights[name]
        w = torch.as_tensor(k).to(dtype).permute(3, 2, 0, 1)        # HWIO -> OIHW
        x = F.relu(F.conv2d(x, w, torch.as_tensor(b).to(dtype), padding=1))
        out[name] = x.permute(0, 2, 3, 1)
    return out


def extractor(image_nhwc, weights, dtype=torch.float64):
    """StyleContentModel.call: {'content': {...}, 'style': {...}}."""
    o = vgg_forward(image_nhwc, weights, dtype)
    return {"content": {n: o[n] for n in CONTENT_LAYERS}, "style": {n: o[n] for n in STYLE_LAYERS}}


def resize_mask(mask_1hw1, size):
    """tf.image.resize default (bilinear, half-pixel centres, antialias=False)  loss.py:112-113."""
    m = mask_1hw1.permute(0, 3, 1, 2)
    if tuple(m.shape[2:]) == tuple(size):
        return mask_1hw1
    m = F.interpolate(m, size=tuple(size), mode="bilinear", align_corners=False, antialias=False)
    return m.permute(0, 2, 3, 1)


def gram_matrix(layer, mask):
    """loss.py:96-102."""
    C = layer.shape[3]
    matrix = layer.reshape(-1, C)
    m = mask.to(layer.dtype).reshape(matrix.shape[0], 1)
    mm = matrix * m
    return mm.t() @ mm


def layer_content_loss(target, output):
    """loss.py:90-92."""
    return torch.mean((target - output) ** 2)


def layer_style_loss(target, output, content_masks, style_masks):
    """loss.py:104-137."""
    out_size = output.shape[1:3]
    tgt_size = target.shape[1:3]
    if content_masks is not None and style_masks is not None:
        sm = [resize_mask(m, tgt_size) for m in style_masks]
        cm = [resize_mask(m, out_size) for m in content_masks]
    else:
        sm = [torch.ones(tuple(tgt_size), dtype=output.dtype)]
        cm = [torch.ones(tuple(out_size), dtype=output.dtype)]
    _, H, W, C = output.shape
    fms, fmc = float(H * W), float(C)
    terms = []
    for c_m, s_m in zip(cm, sm):
        g_t = gram_matrix(output, c_m)
        g_s = gram_matrix(target, s_m)
        mean = torch.mean((g_s - g_t) ** 2)
        terms.append(mean / (2 * fmc ** 2 * fms ** 2))
    return sum(terms)


def iter_on_layers(func, *dicts, **kw):
    """loss.py:80-86: sum over the keys of the first dict, divided by len(args) (== 2, sic)."""
    return sum(func(*[d[k] for d in dicts], **kw) for k in dicts[0].keys()) / len(dicts)


def _sym_index(n, lo, hi):
    """indices of np.pad(arange(n), (lo, hi), 'symmetric')."""
    return torch.as_tensor(np.pad(np.arange(n), (lo, hi), mode="symmetric"))


class V2Torch:
    """matting_v2.MattingLaplacian in torch float64 so autograd differentiates the same graph TF would."""

    def __init__(self, image_hw3, epsilon, window_radius):
        img = torch.as_tensor(image_hw3, dtype=torch.float64)
        self.size = tuple(img.shape)
        self.r = r = int(window_radius)
        H, W, C = self.size
        self.n = n = (2 * r + 1) ** 2
        self.ih, self.iw = _sym_index(H, r + 1, r), _sym_index(W, r + 1, r)
        self.image = self._border(img[..., None])
        iimg = self._ii(self.image)
        prod = self._ii(self.image @ self.image.transpose(-1, -2))
        sums = self._sums(iimg)
        r0 = 2 * r + 1
        sigma = (prod[r0:, r0:] + prod[:H, :W] - prod[r0:, :W] - prod[:H, r0:] - sums @ sums.transpose(-1, -2) / n) / n
        self.means = sums / n
        self.delta_inv = torch.linalg.inv(sigma + (epsilon / n) * torch.eye(C, dtype=torch.float64))

    def _border(self, t):
        return t[self.ih][:, self.iw]

    @staticmethod
    def _ii(t):
        return torch.cumsum(torch.cumsum(t, 0), 1)

    def _sums(self, ii, normalize=False):
        H, W, _ = self.size
        r0 = 2 * self.r + 1
        s = ii[r0:, r0:] + ii[:H, :W] - ii[r0:, :W] - ii[:H, r0:]
        return s / self.n if normalize else s

    def _crop(self, t):
        H, W, _ = self.size
        r = self.r
        return t[r + 1:H + r + 1, r + 1:W + r + 1]

    def matmul(self, x):
        H, W, C = self.size
        p = self._border(x.reshape(H, W, -1, 1))
        p_mean = self._sums(self._ii(p), True)
        ip = self.image @ p.transpose(-1, -2)
        ip_mean = self._sums(self._ii(ip), True)
        a = self.delta_inv @ (ip_mean - self.means @ p_mean.transpose(-1, -2))
        b = p_mean - a.transpose(-1, -2) @ self.means
        a_sum = self._sums(self._ii(self._border(a)))
        b_sum = self._sums(self._ii(self._border(b)))
        q = self.n * self._crop(p) - (a_sum.transpose(-1, -2) @ self._crop(self.image) + b_sum)
        return q.reshape(H * W, -1)


def photorealism(image_nhwc, lap):
    """loss.py:157-161: f64 matvec, result cast back to image dtype."""
    HW = lap.size[0] * lap.size[1]
    p = image_nhwc.reshape(HW, -1).to(torch.float64)
    return torch.sum(p * lap.matmul(p)).to(image_nhwc.dtype)


LOSS_NAMES = {"content": "Content loss", "style": "Style loss", "nima": "NIMA loss",
              "photo": "Photorealism regualarization"}            # loss.py:16-21 (typo is the reference's)
TV_NAME = "Total variation loss"                                   # extension, see total_variation()


def total_variation(image_nhwc):
    """EXTENSION -- NOT IN THE REFERENCE (no TV term anywhere in /root/reference; BASELINE.json's north star
    and configs[1] name one).  Parity unpinned by construction.  Definition = tf.image.total_variation(image)[0]:
    anisotropic L1, sum |x[:,1:,:,:]-x[:,:-1,:,:]| + sum |x[:,:,1:,:]-x[:,:,:-1,:]|, no normalisation; autograd gives
    sign() with sign(0) = 0, as tf.abs does."""
    x = image_nhwc
    return (x[:, 1:] - x[:, :-1]).abs().sum() + (x[:, :, 1:] - x[:, :, :-1]).abs().sum()


def compute_loss(image, outputs, content_target, style_target, weights_cfg, lap=None,
                 content_masks=None, style_masks=None, tv_weight=0.0):
    """loss.py:53-78 with the NIMA term dropped (out of scope, weight must be 0).
    tv_weight > 0 (extension, default off): adds TV_NAME after the reference's terms and tv_weight * TV to the total."""
    vals = {}
    vals["content"] = iter_on_layers(layer_content_loss, content_target, outputs["content"])
    vals["style"] = iter_on_layers(layer_style_loss, style_target, outputs["style"],
                                   content_masks=content_masks, style_masks=style_masks)
    vals["nima"] = torch.zeros((), dtype=image.dtype)
    if weights_cfg["photo"] > 0:
        vals["photo"] = photorealism(image, lap)
    total = sum(weights_cfg[k] * v for k, v in vals.items())
    d = {LOSS_NAMES[k]: v for k, v in vals.items()}
    if tv_weight > 0:
        d[TV_NAME] = total_variation(image)
        total = total + tv_weight * d[TV_NAME]
    d["Total loss"] = total
    return d


def adam_clip_step(x, g, m, v, t, lr=0.1, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.optimizers.Adam.apply_gradients + clip_by_value  (style_transfer.py:321-326,342-343).
    Keras/TF formulation: alpha_t = lr*sqrt(1-b2^t)/(1-b1^t); x -= alpha_t*m/(sqrt(v)+eps)  (eps outside the
    bias correction).  The variable is float32, so TF holds the hyper-parameters as float32 tensors and forms
    (1-beta), beta^t in float32; that is modelled here (it shifts v by 1.3e-5 relative), the state itself stays in
    the dtype of x."""
    f = np.float32
    b1, b2 = f(beta1), f(beta2)
    om1, om2 = float(f(1) - b1), float(f(1) - b2)
    b1p, b2p = f(b1 ** f(t)), f(b2 ** f(t))
    alpha = float(f(lr) * np.sqrt(f(1) - b2p) / (f(1) - b1p))
    m = m + (g - m) * om1
    v = v + (g * g - v) * om2
    x = x - alpha * m / (torch.sqrt(v) + float(f(eps)))
    return torch.clamp(x, 0.0, 1.0), m, v


class TrainState:
    """style_transfer.py:295-344 restated: targets, Laplacian, Adam slots and train_step."""

    def __init__(self, content, style, weights, cfg, content_masks=None, style_masks=None,
                 dtype=torch.float64):
        self.dtype = dtype
        self.weights = weights
        self.cfg = cfg
        self.cm, self.sm = content_masks, style_masks
        with torch.no_grad():
            self.content_target = extractor(content, weights, dtype)["content"]
            self.style_target = extractor(style, weights, dtype)["style"]
        self.lap = None
        if cfg["weights"]["photo"] > 0:
            self.lap = V2Torch(content[0].to(torch.float64), cfg["matting_epsilon"], cfg["matting_window_radius"])
        self.image = content.clone().to(dtype)                      # :329 init = content image
        self.m = torch.zeros_like(self.image)
        self.v = torch.zeros_like(self.image)
        self.t = 0

    def loss_and_grad(self, image=None, return_acts=False):
        """Loss dict and d(Total loss)/d(image) at `image` (default: the current iterate).  return_acts: also the list of
        all 13 conv outputs of this forward pass (detached), for ReLU-decision comparisons (oracle/parity.py)."""
        img = (self.image if image is None else image).clone().requires_grad_(True)
        acts = vgg_forward(img, self.weights, self.dtype)
        outs = {"content": {n: acts[n] for n in CONTENT_LAYERS}, "style": {n: acts[n] for n in STYLE_LAYERS}}
        cm = None if self.cm is None else [m.to(self.dtype) for m in self.cm]
        sm = None if self.sm is None else [m.to(self.dtype) for m in self.sm]
        d = compute_loss(img, outs, self.content_target, self.style_target, self.cfg["weights"],
                         self.lap, cm, sm, tv_weight=self.cfg["weights"].get("tv", 0.0))
        (g,) = torch.autograd.grad(d["Total loss"], img)
        d = {k: float(v.detach()) for k, v in d.items()}
        if return_acts:
            return d, g, [acts[item[0]].detach() for item in VGG_TOPOLOGY if item != "P"]
        return d, g

    def train_step(self):
        d, g = self.loss_and_grad()
        self.t += 1
        a = self.cfg["adam"]
        self.image, self.m, self.v = adam_clip_step(self.image, g, self.m, self.v, self.t,
                                                    a["lr"], a["beta1"], a["beta2"], a["epsilon"])
        return d
