"""Oracle: segmentation-mask helpers (class order).  numpy.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/components/semantic_merge.py:149-170 (get_unique_colors_from_image,
extract_segmentation_masks, mask_for_tf) and :132-138 (reduce_dict).  Pinned against the reference's
own functions through tests/golden/masks_*.npz (oracle/make_golden.py).
"""
import numpy as np


def get_unique_colors_from_image(image):
    """semantic_merge.py:149-154."""
    h, w, c = image.shape
    assert c == 3
    uniq = np.unique(image.reshape(h * w, c), axis=0)
    return [tuple(col) for col in uniq]


def extract_segmentation_masks(segmentation, colors=None):
    """semantic_merge.py:157-165: BGR label image -> {RGB tuple: bool (H,W)}, empty masks dropped."""
    if colors is None:
        colors = [c[::-1] for c in get_unique_colors_from_image(segmentation)]
    out = {}
    seg = segmentation.astype(np.int32)
    for color in colors:
        mask = np.all(seg == color[::-1], axis=-1)
        if mask.max():
            out[color] = mask
    return out


def mask_for_tf(segmentation_mask):
    """semantic_merge.py:168-170: list ordered by sorted(keys) of (1,H,W,1) float32."""
    return [segmentation_mask[k].astype(np.float32)[None, :, :, None] for k in sorted(segmentation_mask)]


def reduce_dict(d, image_shape):
    """semantic_merge.py:132-138: masks -> BGR label image (int)."""
    _, h, w, _ = image_shape
    arr = np.zeros((h, w, 3), int)
    for k, v in d.items():
        I, J = np.where(v)
        arr[I, J] = k[::-1]
    return arr
