"""Deterministic synthetic inputs shared by tests, bench and smoke (SURVEY §8d "Synthetic inputs").

numpy only.  Images are float32 U[0,1) generated directly as floats; label maps are blocky BGR uint8
images from a fixed palette; VGG19 kernels are seeded He-normal, HWIO, float32.
"""
import numpy as np

# (name, Cin, Cout) of the 13 convolutions up to block5_conv1; 'P' marks a 2x2/2 max-pool.
VGG_TOPOLOGY = [
    ("block1_conv1", 3, 64), ("block1_conv2", 64, 64), "P",
    ("block2_conv1", 64, 128), ("block2_conv2", 128, 128), "P",
    ("block3_conv1", 128, 256), ("block3_conv2", 256, 256), ("block3_conv3", 256, 256),
    ("block3_conv4", 256, 256), "P",
    ("block4_conv1", 256, 512), ("block4_conv2", 512, 512), ("block4_conv3", 512, 512),
    ("block4_conv4", 512, 512), "P",
    ("block5_conv1", 512, 512),
]
CONV_LAYERS = [t for t in VGG_TOPOLOGY if t != "P"]

# RGB palette for the label maps (distinct, arbitrary; same for content and style so key sets match).
PALETTE_RGB = [(120, 120, 120), (180, 120, 120), (6, 230, 230), (80, 50, 50), (4, 200, 3),
               (120, 120, 80), (140, 140, 140), (204, 5, 255), (230, 230, 230), (4, 250, 7)]


def image(H, W, seed):
    """(1,H,W,3) float32 U[0,1)."""
    return np.random.default_rng(seed).random((1, H, W, 3), dtype=np.float32)


def smooth_image(H, W, seed, passes=6):
    """(1,H,W,3) float32 low-pass-filtered noise quantised to k/255 (photo-like conditioning)."""
    a = np.random.default_rng(seed).random((H, W, 3))
    for _ in range(passes):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5.0
    a = (a - a.min()) / (a.max() - a.min())
    return (np.round(a * 255.0) / 255.0).astype(np.float32)[None]


def label_image(H, W, K, seed, cell=32):
    """(H,W,3) uint8 BGR label image with exactly K classes laid out in cell x cell blocks."""
    assert 1 <= K <= len(PALETTE_RGB)
    gh, gw = -(-H // cell), -(-W // cell)
    assert gh * gw >= K, "image too small for K classes at this cell size"
    rng = np.random.default_rng(seed)
    grid = rng.integers(0, K, size=gh * gw)
    grid[rng.permutation(gh * gw)[:K]] = np.arange(K)       # every class present
    grid = grid.reshape(gh, gw)
    lab = np.kron(grid, np.ones((cell, cell), dtype=np.int64))[:H, :W]
    pal_bgr = np.array([c[::-1] for c in PALETTE_RGB[:K]], dtype=np.uint8)
    return pal_bgr[lab]


def vgg_weights(seed=1234, bias_scale=1.0):
    """dict name -> (kernel (3,3,Cin,Cout) f32 He-normal, bias (Cout,) f32)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, cin, cout in CONV_LAYERS:
        std = np.sqrt(2.0 / (9 * cin))
        k = (rng.standard_normal((3, 3, cin, cout)) * std).astype(np.float32)
        b = (rng.standard_normal(cout) * bias_scale).astype(np.float32)
        out[name] = (k, b)
    return out
