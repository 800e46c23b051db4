"""Mask helpers of /root/reference/components/semantic_merge.py that touch the hot path:
they fix the CLASS ORDER of the masked Gram loss.  Host-side numpy, run once before the loop.

Provided: get_unique_colors_from_image (:149-154), extract_segmentation_masks (:157-165), mask_for_tf (:168-170),
reduce_dict (:132-138), replace_colors_in_dict (:27-35).  The WordNet-based merge_segments (:21-129) needs sematch,
nltk corpora and a PSPNet label pickle that are not in the tree; it is out of scope and raises.
"""
import numpy as np
import torch


def get_unique_colors_from_image(image):
    """Distinct colours of an (H,W,3) uint8 image as tuples, in lexicographic order of (c0,c1,c2) -- the order
    np.unique(axis=0) gives in the reference."""
    h, w, c = image.shape
    assert c == 3
    v = image.reshape(h * w, 3).astype(np.int64)
    packed = np.unique((v[:, 0] << 16) | (v[:, 1] << 8) | v[:, 2])
    return [(image.dtype.type(p >> 16), image.dtype.type((p >> 8) & 255), image.dtype.type(p & 255)) for p in packed]


def extract_segmentation_masks(segmentation, colors=None):
    """BGR label image (cv2.imread order) -> {RGB tuple: bool mask (H,W)}; empty masks are dropped."""
    if colors is None:
        colors = [color[::-1] for color in get_unique_colors_from_image(segmentation)]
    seg = segmentation.astype(np.int64)
    packed = (seg[:, :, 0] << 16) | (seg[:, :, 1] << 8) | seg[:, :, 2]      # one comparison per colour instead of three
    out = {}
    for color in colors:
        b, g, r = (int(c) for c in color[::-1])
        if min(b, g, r) < 0 or max(b, g, r) > 255:
            continue                                   # cannot occur in a uint8 image
        mask = packed == ((b << 16) | (g << 8) | r)
        if mask.any():
            out[color] = mask
    return out


def mask_for_tf(segmentation_mask, device=None):
    """List ordered by sorted(keys) of (1,H,W,1) float32 tensors (the reference returns tf constants)."""
    return [torch.as_tensor(segmentation_mask[key].astype(np.float32))[None, :, :, None].to(device or "cpu")
            for key in sorted(segmentation_mask)]


def reduce_dict(dict, image):
    """Masks -> BGR label image (int array), the inverse of extract_segmentation_masks."""
    _, h, w, _ = image.shape
    arr = np.zeros((h, w, 3), int)
    for k, v in dict.items():
        arr[v] = k[::-1]
    return arr


def replace_colors_in_dict(color_mask_dict, replacement_colors):
    """Rename colours, OR-ing masks that collapse onto the same colour."""
    merged = {}
    for color, mask in color_mask_dict.items():
        new = replacement_colors.get(color, color)
        merged[new] = np.logical_or(mask, merged[new]) if new in merged else mask
    return merged


def merge_segments(*_args, **_kwargs):
    raise NotImplementedError("semantic merging (semantic_merge.py:71-129) needs sematch/WordNet and the PSPNet label "
                              "table, which are outside the B200 hot path; supply merged *_seg.png files instead")
