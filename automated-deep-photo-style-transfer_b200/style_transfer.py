"""Hot loop of /root/reference/style_transfer.py (:295-367) on libadpst.

The reference inlines everything in `if __name__ == "__main__"`.  Here the same steps are callable:
    extractor = StyleContentModel(content_layers, style_layers, weights=...)       # :295-298
    loss      = Loss(content_target, style_target, args, content_masks, style_masks)# :304-310
    loss.initialize_matting_laplacian(content[0].double())                          # :312-317
    opt       = Adam(learning_rate, beta_1, beta_2, epsilon)                        # :321-326
    step      = make_train_step(extractor, loss, opt)                               # :331-344
    for i in range(iters): loss_dict = step(transfer_image)                         # :353-367
`style_transfer(...)` wraps exactly that and returns the best image like the reference's loop (:366-367).
The flags of the reference's argparse block (:126-205) are reproduced by `build_parser()` and `main()` follows the script
(:207-384): meta.json (:97-120), *_seg.png mask files (:223-264), best-image tracking (:366-367), loss printing (:360-364),
per-iteration scalars and intermediate images (the reference's TensorBoard summaries, :372-381, as files).
"""
import argparse
import json
import os
import time
from pathlib import Path

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


_SIDE_STREAMS = {}


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


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
            s = _side_stream(image.device)
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
            # Capture with the raw CUDAGraph API: the torch.cuda.graph() context manager runs gc.collect() and
            # torch.cuda.empty_cache() on entry, which hands every cached block back to the driver -- measured 0.3 s per
            # capture plus a cudaMalloc for every buffer of the next pair.  The step allocates nothing while it is captured
            # (every buffer was created by the warm-up), so the graph's private pool stays empty.
            g = torch.cuda.CUDAGraph()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                g.capture_begin()
                try:
                    out = eager_step(image)
                finally:
                    g.capture_end()
            torch.cuda.current_stream().wait_stream(s)
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
    """style_transfer.py:69-79: uint8(255*x) by truncation (tf.cast float -> uint8), batch dimension squeezed.
    Returns a (H,W,3) uint8 tensor (the reference PNG-encodes it; `save_image` below does that on the host)."""
    return (255 * tensor).to(torch.uint8).squeeze(0)


def load_image(filename, dtype="float32"):
    """style_transfer.py:50-59: decode a JPEG/PNG to RGB, convert_image_dtype (uint8 -> float: multiply by 1/255 in the
    target type), add the batch dimension.  Returns a (1,H,W,3) numpy array.  (cv2/libjpeg-turbo stands in for TensorFlow's
    decoder, which is not installable here: decoded JPEG pixels may differ by +-1 between decoders.)"""
    import cv2
    import numpy as np
    bgr = cv2.imread(str(filename), cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError("Image file {} does not exist.".format(filename))
    dt = np.dtype(dtype)
    return (bgr[:, :, ::-1].astype(dt) * dt.type(1.0 / 255.0))[None]


def save_image(image_u8, file):
    """style_transfer.py:61-62 (+ encode_png of :79): write a (H,W,3) uint8 RGB tensor/array; the extension selects the codec."""
    import cv2
    import numpy as np
    arr = image_u8.cpu().numpy() if isinstance(image_u8, torch.Tensor) else np.asarray(image_u8)
    if not cv2.imwrite(str(file), np.ascontiguousarray(arr[:, :, ::-1])):
        raise OSError("could not write %s" % file)


def change_filename(dir_name, filename, suffix, extension=None):
    r"""style_transfer.py:82-94.  change_filename('.', 'image.png', '_seg') -> ./image_seg.png"""
    path, ext = os.path.splitext(filename)
    if extension is None:
        extension = ext
    return os.path.join(dir_name, path + suffix + extension)


def write_metadata(args, load_segmentation, experiment_path):
    """style_transfer.py:97-120: the same keys, in the same order, into <experiment_path>/meta.json."""
    meta = {
        "init": args.init,
        "iter": args.iter,
        "content": args.content_image,
        "style": args.style_image,
        "content_weight": args.content_weight,
        "style_weight": args.style_weight,
        "regularization_weight": args.regularization_weight,
        "nima_weight": args.nima_weight,
        "semantic_thresh": args.semantic_thresh,
        "similarity_metric": args.similarity_metric,
        "load_segmentation": load_segmentation,
        "adam": {
            "learning_rate": args.adam_lr,
            "beta1": args.adam_beta1,
            "beta2": args.adam_beta2,
            "epsilon": args.adam_epsilon
        }
    }
    file = Path(experiment_path) / 'meta.json'
    with file.open('w+') as f:
        f.write(json.dumps(meta, indent=4))
    return meta


def build_parser():
    """The argument groups and flags of style_transfer.py:126-205 with the reference's defaults, except
    --nima_weight (reference default 1e5: the NIMA term is outside this hot path, so the default is 0 and any other value is
    refused with an explanation).  Flags after the 'Extensions' header do not exist in the reference."""
    parser = argparse.ArgumentParser(description="B200 hot path of automated deep photo style transfer")
    base = parser.add_argument_group('Base options')
    expr = parser.add_argument_group('Experiment parameters')
    param = parser.add_argument_group('Hyperparameters')
    dirs = parser.add_argument_group('Storage directories')
    misc = parser.add_argument_group('Miscellaneous')
    ext = parser.add_argument_group('Extensions (not in the reference)')

    base.add_argument("-c", "--content_image", type=str, help="Content image path", default="blanc.jpg")
    base.add_argument("-s", "--style_image", type=str, help="Style image path", default="bear.jpeg")
    base.add_argument("-o", "--output_image", type=str, help="Output image path, default: result.jpg", default="result.jpg")

    expr.add_argument("--dtype", type=str, help="dtype of the input and output images., default: float32", default="float32")
    expr.add_argument("--init", type=str, help="Initialization image., default: content",
                      choices=["noise", "content", "style"], default="content")
    expr.add_argument("--iter", type=int, help="Number of iterations", default=1000)
    expr.add_argument("--similarity_metric", type=str, help="Semantic similarity metric for label grouping., default: li",
                      choices=["li", "wpath", "jcn", "lin", "wup", "res"], default="li")

    param.add_argument("--content_weight", type=float, help="Weight of the content loss., default: 1", default=1)
    param.add_argument("--style_weight", type=float, help="Weight of the style loss., default: 100", default=1e2)
    param.add_argument("--regularization_weight", type=float, help="Weight of the photorealism regularization.", default=1e4)
    param.add_argument("--nima_weight", type=float, default=0,
                       help="Weight for nima loss. Reference default 1e5; the NIMA term is not built here: must be 0")
    param.add_argument("--adam_lr", type=float, help="Learning rate for the adam optimizer.", default=1e-1)
    param.add_argument("--adam_beta1", type=float, help="Beta1 for the adam optimizer., default: 0.9", default=0.9)
    param.add_argument("--adam_beta2", type=float, help="Beta2 for the adam optimizer., default: 0.999", default=0.999)
    param.add_argument("--adam_epsilon", type=float, help="Epsilon for the adam optimizer., default: 1e-08", default=1e-08)
    param.add_argument("--matting_epsilon", type=float, help="Epsilon regularization for matting laplacian computing.",
                       default=1e-5)
    param.add_argument("--matting_window_radius", type=int, help="Size of the windows considered by matting laplacian.",
                       default=3)
    param.add_argument("--semantic_thresh", type=float, help="Semantic threshold for label grouping., default: 0.5", default=0.5)

    dirs.add_argument("--logs_dir", type=Path, help="Path to the per-iteration logs., default: ./logs", default=Path('logs'))
    dirs.add_argument("--results_dir", type=Path, help='Where results are stored., default: ./experiments',
                      default=Path('experiments'))
    dirs.add_argument("--seg_dir", type=Path, help='Where segmented images are stored., default: ./raw_seg',
                      default=Path('raw_seg'))

    misc.add_argument("--gpu", type=str, help="Comma separated list of GPU(s) to use.", default="0")
    misc.add_argument("--experiment_name", type=str, help="Name of the experiment., default: <timestamp>", default=None)
    misc.add_argument("--intermediate_result_interval", type=int,
                      help="Interval of iterations until a intermediate result is saved.", default=20)
    misc.add_argument("--print_loss_interval", type=int,
                      help="Interval of iterations until the current loss is printed to console.", default=10)

    ext.add_argument('--vgg_weights', type=str, default=None,
                     help=".npz with '<layer>/kernel' (3,3,Cin,Cout) and '<layer>/bias' (Keras downloads ImageNet weights, "
                          "VGG19/model.py:7; offline they must be supplied, see also ADPST_VGG19_WEIGHTS)")
    ext.add_argument('--matting', type=str, default='v2', choices=['v2', 'v3'],
                     help="Laplacian variant (the reference hard-wires v2, loss.py:4)")
    ext.add_argument('--use_masks', action='store_true',
                     help="pass the segmentation masks to Loss (commented out in the reference script, :308-309)")
    ext.add_argument('--tv_weight', type=float, default=0.0, help="total-variation weight (no such term in the reference)")
    return parser


def main(argv=None):
    """python -m ... .style_transfer [flags]: the reference script, :207-384, on the B200 hot path."""
    import cv2
    import numpy as np
    from .components.semantic_merge import extract_segmentation_masks, mask_for_tf, reduce_dict
    args = build_parser().parse_args(argv)
    if args.nima_weight != 0:
        raise SystemExit("--nima_weight %g: the NIMA term (components/loss.py:141-153, InceptionResNetV2 with weights that are "
                         "not in the tree) is not part of this build; run with --nima_weight 0 (also in the reference, for "
                         "identical objectives)" % args.nima_weight)
    if args.gpu:                                                               # :210-211
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    if not args.experiment_name:                                               # :217-219
        from datetime import datetime
        args.experiment_name = datetime.now().strftime('%Y-%m-%d_%H:%M')
    experiment_path = args.results_dir / args.experiment_name
    experiment_path.mkdir(parents=True, exist_ok=True)
    args.seg_dir.mkdir(parents=True, exist_ok=True)

    # manual segmentation masks (:224-228)
    content_segmentation_filename = change_filename(args.seg_dir, args.content_image, '_seg', '.png')
    style_segmentation_filename = change_filename(args.seg_dir, args.style_image, '_seg', '.png')
    load_segmentation = os.path.exists(content_segmentation_filename) and os.path.exists(style_segmentation_filename)
    write_metadata(args, load_segmentation, experiment_path)

    for file in [args.content_image, args.style_image]:                       # :234-237
        if not os.path.exists(file):
            print("Image file {} does not exist.".format(file))
            raise SystemExit(1)
    content_image = load_image(args.content_image, args.dtype)               # :241-242
    style_image = load_image(args.style_image, args.dtype)

    content_masks = style_masks = None
    if load_segmentation:                                                      # :245-250
        print("Load segmentation from files.")
        content_segmentation_masks = extract_segmentation_masks(cv2.imread(content_segmentation_filename))
        style_segmentation_masks = extract_segmentation_masks(cv2.imread(style_segmentation_filename))
        os.makedirs(os.path.dirname(content_segmentation_filename) or ".", exist_ok=True)
        cv2.imwrite(content_segmentation_filename, reduce_dict(content_segmentation_masks, content_image))   # :261-264
        cv2.imwrite(style_segmentation_filename, reduce_dict(style_segmentation_masks, style_image))
        if args.use_masks:
            if set(content_segmentation_masks) != set(style_segmentation_masks):
                raise SystemExit("content and style segmentations use different colour sets (semantic_merge.py:127 asserts "
                                 "equal key sets after merging)")
            content_masks, style_masks = mask_for_tf(content_segmentation_masks), mask_for_tf(style_segmentation_masks)
    else:
        # :252-259 would run PSPNet + the WordNet merge; both are outside this hot path (weights / corpora absent)
        print("No *_seg.png files in {}: running without segmentation masks (PSPNet segmentation is not part of this build)."
              .format(args.seg_dir))
        if args.use_masks:
            raise SystemExit("--use_masks needs {} and {}".format(content_segmentation_filename, style_segmentation_filename))
    if args.init != "content":                                                # :266-278; the reference ignores init (:329)
        print("note: --init {} is parsed but, as in the reference (style_transfer.py:329), the variable starts from the "
              "content image".format(args.init))

    iterations_dir = experiment_path / 'iter'
    iterations_dir.mkdir(exist_ok=True)
    logs_path = args.logs_dir / args.experiment_name
    logs_path.mkdir(parents=True, exist_ok=True)

    print("Style transfer started")
    scalars = (logs_path / 'scalars.jsonl').open('w')                          # stands in for tf.summary.scalar (:372-376)
    t_loop = time.time()

    def on_iteration(i, losses, image):
        if i % args.print_loss_interval == 0:                                  # :360-364
            print("[Iter {}]".format(i), end='\t')
            for loss_name, loss_value in losses.items():
                print('{}: {:<15.3f}'.format(loss_name, loss_value), end='')
            print()
        scalars.write(json.dumps(dict(losses, step=i)) + "\n")
        if i % args.intermediate_result_interval == 0:                         # :379-381 tf.summary.image
            save_image(tensor_to_image(image), iterations_dir / "iter_{}.png".format(i))

    best, history = style_transfer(content_image, style_image, args, content_masks, style_masks, vgg_weights=args.vgg_weights,
                                   matting=args.matting, callback=on_iteration)
    scalars.close()
    print("Style transfer finished. Average time per epoch: {:.5f}s\n".format((time.time() - t_loop) / max(args.iter, 1)))
    if args.output_image and best is not None:
        out = experiment_path / os.path.basename(args.output_image)
        save_image(tensor_to_image(best), out)
        print("Best image (lowest total loss) written to {}".format(out))
    return best, history


if __name__ == "__main__":
    main()
