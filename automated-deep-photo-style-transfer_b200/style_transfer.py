"""Hot loop of /root/reference/style_transfer.py (:295-367) on libadpst.

The reference inlines everything in `if __name__ == "__main__"`.  Here the same steps are callable:
    extractor = StyleContentModel(content_layers, style_layers, weights=...)       # :295-298
    loss      = Loss(content_target, style_target, args, content_masks, style_masks)# :304-310
    loss.initialize_matting_laplacian(content[0].double())                          # :312-317
    opt       = Adam(learning_rate, beta_1, beta_2, epsilon)                        # :321-326
    step      = make_train_step(extractor, loss, opt)                               # :331-344
    for i in range(iters): loss_dict = step(transfer_image)                         # :353-367
`style_transfer(...)` wraps exactly that and returns the best image like the reference's loop (:366-367).
The flags of the reference's argparse block (:126-205) are reproduced by `build_parser()`.
"""
import argparse
import time

import torch

from . import kernels
from .components.VGG19.model import StyleContentModel
from .components.loss import Loss

CONTENT_LAYERS = ['block4_conv2']                                              # style_transfer.py:295
STYLE_LAYERS = ['block%d_conv1' % (i + 1) for i in range(5)]                   # style_transfer.py:296


class Adam:
    """tf.optimizers.Adam(learning_rate, beta_1, beta_2, epsilon) for one image variable, with the clip of
    style_transfer.py:343 fused into the update kernel."""

    def __init__(self, learning_rate=0.1, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self._slots = None

    def apply_gradients_and_clip(self, grad, image):
        if self._slots is None or self._slots.m.shape != image.shape:
            self._slots = kernels.AdamState(image)
        kernels.adam_clip_step(image, grad, self._slots, self.learning_rate, self.beta_1, self.beta_2, self.epsilon)

    @property
    def iterations(self):
        return 0 if self._slots is None else self._slots.step


def make_train_step(features_extractor, compute_loss, optimizer, use_cuda_graph=False):
    """Returns train_step(image) -> loss_dict, the closure of style_transfer.py:331-344.
    `image` is a (1,H,W,3) float32 CUDA tensor updated in place (the tf.Variable of :329).
    With use_cuda_graph the whole step is captured once and replayed (no per-launch host cost)."""
    grad_buf = {}

    def eager_step(image):
        outputs = features_extractor(image, reuse=True)                        # :335
        loss_dict = compute_loss(image, outputs)                               # :336
        g = grad_buf.get("g")
        if g is None or g.shape != image.shape:
            g = grad_buf["g"] = torch.empty_like(image)
        grad = compute_loss.gradient(features_extractor, out=g)                # :341
        optimizer.apply_gradients_and_clip(grad, image)                        # :342-343
        return loss_dict                                                       # :344

    if not use_cuda_graph:
        return eager_step

    state = {"graph": None, "image": None, "out": None}

    def graphed_step(image):
        if state["graph"] is None or state["image"] is not image:
            # warm up on a side stream (allocations, lazy state), restoring the variable afterwards
            snap = image.clone()
            slots = optimizer._slots
            saved = None if slots is None else (slots.m.clone(), slots.v.clone(), slots.state.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                eager_step(image)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            image.copy_(snap)
            if saved is None:
                optimizer._slots.m.zero_(); optimizer._slots.v.zero_(); optimizer._slots.state.zero_()
            else:
                slots.m.copy_(saved[0]); slots.v.copy_(saved[1]); slots.state.copy_(saved[2])
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = eager_step(image)
            state.update(graph=g, image=image, out=out)
        state["graph"].replay()
        return state["out"]

    return graphed_step


def style_transfer(content_image, style_image, args, content_masks=None, style_masks=None, vgg_weights=None,
                   matting="v2", use_cuda_graph=True, callback=None):
    """content_image / style_image: (1,H,W,3) float32 in [0,1] (CUDA or host).  args: namespace with the
    reference's hyper-parameter flags.  Returns (best_image (1,H,W,3) float32 CUDA tensor, history of loss dicts)."""
    dev = torch.device("cuda")
    content_image = torch.as_tensor(content_image, dtype=torch.float32).to(dev).contiguous()
    style_image = torch.as_tensor(style_image, dtype=torch.float32).to(dev).contiguous()

    features_extractor = StyleContentModel(CONTENT_LAYERS, STYLE_LAYERS, shape=(None, None, 3), weights=vgg_weights)
    content_target = features_extractor(content_image)['content']             # :301
    style_target = features_extractor(style_image)['style']                   # :302
    compute_loss = Loss(content_target, style_target, args, content_masks, style_masks, matting=matting)   # :304-310
    if args.regularization_weight > 0:                                         # :312-317
        compute_loss.initialize_matting_laplacian(content_image[0].to(torch.float64))
    optimizer = Adam(args.adam_lr, args.adam_beta1, args.adam_beta2, args.adam_epsilon)     # :321-326
    transfer_image = content_image.clone()                                     # :329 (init is always the content image)
    train_step = make_train_step(features_extractor, compute_loss, optimizer, use_cuda_graph)

    min_loss, best_image, history = float('inf'), None, []
    for i in range(1, args.iter + 1):                                          # :353
        loss_dict = train_step(transfer_image)
        host = {k: float(v) for k, v in loss_dict.items()}                     # one sync per iteration, as :366
        history.append(host)
        # the reference evaluates the loss before the update and snapshots the image after it (:356, :366-367)
        if host['Total loss'] < min_loss:
            min_loss, best_image = host['Total loss'], transfer_image.clone()
        if callback is not None:
            callback(i, host, transfer_image)
    return best_image, history


def tensor_to_image(tensor):
    """style_transfer.py:69-79: uint8(255*x) by truncation, batch dimension dropped.  Returns a (H,W,3) uint8 tensor."""
    return (255 * tensor).to(torch.uint8).squeeze(0)


def build_parser():
    """The hyper-parameter / experiment flags of style_transfer.py:126-205 with the reference's defaults."""
    p = argparse.ArgumentParser(description="B200 hot path of automated deep photo style transfer")
    p.add_argument('-c', '--content_image', type=str, default='blanc.jpg')
    p.add_argument('-s', '--style_image', type=str, default='bear.jpeg')
    p.add_argument('-o', '--output_image', type=str, default=None)
    p.add_argument('--dtype', type=str, default='float32')
    p.add_argument('--init', type=str, default='content', choices=['noise', 'content', 'style'])
    p.add_argument('--iter', type=int, default=1000)
    p.add_argument('--content_weight', type=float, default=1)
    p.add_argument('--style_weight', type=float, default=100)
    p.add_argument('--regularization_weight', type=float, default=10 ** 4)
    p.add_argument('--nima_weight', type=float, default=0, help="reference default 1e5; NIMA is out of scope, must be 0")
    p.add_argument('--adam_lr', type=float, default=0.1)
    p.add_argument('--adam_beta1', type=float, default=0.9)
    p.add_argument('--adam_beta2', type=float, default=0.999)
    p.add_argument('--adam_epsilon', type=float, default=1e-08)
    p.add_argument('--matting_epsilon', type=float, default=1e-5)
    p.add_argument('--matting_window_radius', type=int, default=3)
    p.add_argument('--print_loss_interval', type=int, default=1)
    p.add_argument('--vgg_weights', type=str, default=None, help=".npz with '<layer>/kernel' and '<layer>/bias'")
    p.add_argument('--matting', type=str, default='v2', choices=['v2', 'v3'])
    return p


def main(argv=None):
    import cv2
    import numpy as np
    args = build_parser().parse_args(argv)

    def load_image(fn):                                                        # style_transfer.py:55-64
        bgr = cv2.imread(fn, cv2.IMREAD_COLOR)
        if bgr is None:
            raise SystemExit("Image file {} does not exist.".format(fn))
        return (bgr[:, :, ::-1].astype(np.float32) * np.float32(1.0 / 255.0))[None]

    t0 = time.time()

    def show(i, d, _img):
        if i % args.print_loss_interval == 0:
            print("[Iter {}]".format(i), end='\t')
            for k, v in d.items():
                print('{}: {:<15.3f}'.format(k, v), end='')
            print()

    best, hist = style_transfer(load_image(args.content_image), load_image(args.style_image), args,
                                vgg_weights=args.vgg_weights, matting=args.matting, callback=show)
    print("Style transfer finished. Average time per epoch: {:.5f}s\n".format((time.time() - t0) / max(args.iter, 1)))
    if args.output_image:
        cv2.imwrite(args.output_image, tensor_to_image(best).cpu().numpy()[:, :, ::-1])


if __name__ == "__main__":
    main()
