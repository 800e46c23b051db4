"""Oracle: matting-Laplacian operators, numpy float64.  TEST INFRASTRUCTURE ONLY.

v2 ("large kernel", He et al.) follows /root/reference/components/matting_v2.py line by line;
v3 (explicit Levin COO) follows /root/reference/components/matting_v3.py:61-100.
tf.pad(mode='SYMMETRIC') == np.pad(mode='symmetric'); tf.cumsum == np.cumsum.
"""
import numpy as np
import scipy.sparse


# --------------------------------------------------------------------------------------
# v2: matrix-free operator, integral images, symmetric padding (matting_v2.py)
# --------------------------------------------------------------------------------------
def _t(a):
    """matting_v2.py:227-229  _transpose: swap the two trailing (matrix) axes."""
    return np.swapaxes(a, -1, -2)


def add_border(img, r, mode="symmetric"):
    """matting_v2.py:180-197: pad H and W by [r+1, r]."""
    pad = [(r + 1, r), (r + 1, r)] + [(0, 0)] * (img.ndim - 2)
    if mode == "symmetric":
        return np.pad(img, pad, mode="symmetric")
    return np.pad(img, pad, mode="constant")


def crop_border(img, H, W, r):
    """matting_v2.py:199-202."""
    return img[r + 1:H + r + 1, r + 1:W + r + 1]


def integral_image(img):
    """matting_v2.py:231-251: cumsum along axis 0 then axis 1."""
    return np.cumsum(np.cumsum(img, axis=0), axis=1)


def compute_sums(iimg, H, W, r, normalize=False):
    """matting_v2.py:205-217: four-corner difference of the integral image."""
    n = (2 * r + 1) ** 2
    r0 = c0 = 2 * r + 1
    r1, c1 = H, W
    sums = iimg[r0:, c0:] + iimg[:r1, :c1] - iimg[r0:, :c1] - iimg[:r1, c0:]
    return sums / n if normalize else sums


def compute_covs(Q, sums, H, W, r):
    """matting_v2.py:219-225."""
    n = (2 * r + 1) ** 2
    r0 = c0 = 2 * r + 1
    r1, c1 = H, W
    return (Q[r0:, c0:] + Q[:r1, :c1] - Q[r0:, :c1] - Q[:r1, c0:] - (sums @ _t(sums)) / n) / n


class V2Operator:
    """State of matting_v2.MattingLaplacian.__init__ (matting_v2.py:11-52)."""

    def __init__(self, image, epsilon=1e-5, window_radius=1):
        image = np.asarray(image)
        self.dtype = image.dtype
        self.radius = r = int(window_radius)
        self.size = image.shape
        H, W, C = image.shape
        self.image = add_border(image[..., None], r)                 # :36
        iimg = integral_image(self.image)                            # :39
        prod_image = self.image @ _t(self.image)                     # :40
        prod_iimg = integral_image(prod_image)                       # :41
        n = self.window_area = (2 * r + 1) ** 2                      # :44
        sums = compute_sums(iimg, H, W, r)                           # :49
        sigma = compute_covs(prod_iimg, sums, H, W, r)               # :50
        self.means = sums / n                                        # :51
        eye = np.eye(C, dtype=image.dtype)
        self.delta_inv = np.linalg.inv(sigma + (epsilon / n) * eye)  # :52

    @property
    def shape(self):
        H, W, _ = self.size
        return (H * W, H * W)

    def matmul(self, x):
        """matting_v2.py:147-176  _matmul."""
        H, W, C = self.size
        r = self.radius
        x = np.asarray(x, dtype=self.dtype)
        p = add_border(x.reshape(H, W, -1, 1), r)                                   # :153
        p_mean = compute_sums(integral_image(p), H, W, r, normalize=True)           # :154
        ip = self.image @ _t(p)                                                     # :157
        ip_mean = compute_sums(integral_image(ip), H, W, r, normalize=True)         # :159
        a_star = self.delta_inv @ (ip_mean - self.means @ _t(p_mean))               # :161
        b_star = p_mean - _t(a_star) @ self.means                                   # :162
        a_sum = compute_sums(integral_image(add_border(a_star, r)), H, W, r)        # :164
        b_sum = compute_sums(integral_image(add_border(b_star, r)), H, W, r)        # :165
        q = self.window_area * crop_border(p, H, W, r) - \
            (_t(a_sum) @ crop_border(self.image, H, W, r) + b_sum)                  # :167-168
        return q.reshape(H * W, -1)                                                 # :176


def v2_direct(image, x, epsilon, r):
    """Independent second restatement of the v2 operator: per-window loops, no
    integral images.  O(HW (2r+1)^2); small inputs only.  Used to cross-check V2Operator."""
    image = np.asarray(image, np.float64)
    H, W, C = image.shape
    x = np.asarray(x, np.float64).reshape(H, W, -1)
    Cx = x.shape[-1]
    n = (2 * r + 1) ** 2
    It = np.pad(image, [(r, r), (r, r), (0, 0)], mode="symmetric")
    xt = np.pad(x, [(r, r), (r, r), (0, 0)], mode="symmetric")
    a = np.zeros((H, W, C, Cx))
    b = np.zeros((H, W, Cx))
    for i in range(H):
        for j in range(W):
            wi = It[i:i + 2 * r + 1, j:j + 2 * r + 1].reshape(n, C)
            wx = xt[i:i + 2 * r + 1, j:j + 2 * r + 1].reshape(n, Cx)
            mu = wi.mean(0)
            pb = wx.mean(0)
            sig = (wi.T @ wi - n * np.outer(mu, mu)) / n
            dinv = np.linalg.inv(sig + (epsilon / n) * np.eye(C))
            a[i, j] = dinv @ (wi.T @ wx / n - np.outer(mu, pb))
            b[i, j] = pb - a[i, j].T @ mu
    at = np.pad(a, [(r, r), (r, r), (0, 0), (0, 0)], mode="symmetric")
    bt = np.pad(b, [(r, r), (r, r), (0, 0)], mode="symmetric")
    y = np.zeros((H, W, Cx))
    for i in range(H):
        for j in range(W):
            asum = at[i:i + 2 * r + 1, j:j + 2 * r + 1].sum((0, 1))
            bsum = bt[i:i + 2 * r + 1, j:j + 2 * r + 1].sum((0, 1))
            y[i, j] = n * x[i, j] - (asum.T @ image[i, j] + bsum)
    return y.reshape(H * W, Cx)


def _chol3_solve_ld(M, R):
    """Solve M A = R for symmetric positive definite 3x3 M in np.longdouble by Cholesky (np.linalg has no longdouble)."""
    l00 = np.sqrt(M[0, 0]); l10 = M[0, 1] / l00; l20 = M[0, 2] / l00
    l11 = np.sqrt(M[1, 1] - l10 * l10); l21 = (M[1, 2] - l20 * l10) / l11
    l22 = np.sqrt(M[2, 2] - l20 * l20 - l21 * l21)
    z0 = R[0] / l00; z1 = (R[1] - l10 * z0) / l11; z2 = (R[2] - l20 * z0 - l21 * z1) / l22
    a2 = z2 / l22; a1 = (z1 - l21 * a2) / l11; a0 = (z0 - l10 * a1 - l20 * a2) / l00
    return np.stack([a0, a1, a2])


def v2_extended_precision(image, x, epsilon, r=1):
    """The v2 operator (same definition as v2_direct / matting_v2.py) evaluated in np.longdouble (64-bit mantissa on x86) with
    centred window moments and Cholesky solves.  It exists because the reference's own float64 arithmetic is NOT accurate
    where the result is a tiny remainder: at x = I (iteration 0) |L I| ~ 1e-7 and integral images plus explicit 3x3 inverses
    leave ~1e-11 of absolute noise (4.7e-5 relative on I^T L I for a grey 2x3 image), more than the GPU kernel's error, so
    that case is judged against this restatement.  O(HW (2r+1)^2) Python loops: small inputs only."""
    LD = np.longdouble
    image = np.asarray(image, LD)
    H, W, C = image.shape
    x = np.asarray(x, LD).reshape(H, W, -1)
    Cx = x.shape[-1]
    d = 2 * r + 1
    n = d * d
    It = np.pad(image, [(r, r), (r, r), (0, 0)], mode="symmetric")
    xt = np.pad(x, [(r, r), (r, r), (0, 0)], mode="symmetric")
    a = np.zeros((H, W, C, Cx), LD)
    b = np.zeros((H, W, Cx), LD)
    for i in range(H):
        for j in range(W):
            wi = It[i:i + d, j:j + d].reshape(n, C)
            wx = xt[i:i + d, j:j + d].reshape(n, Cx)
            mu, pb = wi.mean(0), wx.mean(0)
            ci, cx = wi - mu, wx - pb
            a[i, j] = _chol3_solve_ld(ci.T @ ci + LD(epsilon) * np.eye(C, dtype=LD), ci.T @ cx)
            b[i, j] = pb - a[i, j].T @ mu
    at = np.pad(a, [(r, r), (r, r), (0, 0), (0, 0)], mode="symmetric")
    bt = np.pad(b, [(r, r), (r, r), (0, 0)], mode="symmetric")
    y = np.zeros((H, W, Cx), LD)
    for i in range(H):
        for j in range(W):
            y[i, j] = n * x[i, j] - (at[i:i + d, j:j + d].sum((0, 1)).T @ image[i, j] + bt[i:i + d, j:j + d].sum((0, 1)))
    return y.reshape(H * W, Cx)


# --------------------------------------------------------------------------------------
# v3: explicit COO Laplacian, interior windows only (matting_v3.py:61-100)
# --------------------------------------------------------------------------------------
def v3_compute_laplacian(img, eps=1e-5, win_rad=1):
    """Returns (rows, cols, vals, shape) in the reference's emission order; duplicates kept."""
    img = np.asarray(img)
    win_size = (win_rad * 2 + 1) ** 2                       # :74
    h, w, d = img.shape
    c_h, c_w = h - 2 * win_rad, w - 2 * win_rad             # :77
    win_diam = win_rad * 2 + 1
    indsM = np.arange(h * w).reshape((h, w))                # :80
    ravelImg = img.reshape(h * w, d)                        # :81
    win_inds = np.lib.stride_tricks.sliding_window_view(indsM, (win_diam, win_diam))   # :82 _rolling_block
    win_inds = win_inds.reshape(c_h * c_w, win_size)        # :84-86
    winI = ravelImg[win_inds]                               # :87
    win_mu = np.mean(winI, axis=1, keepdims=True)           # :89
    win_var = np.einsum('...ji,...jk ->...ik', winI, winI) / win_size - \
        np.einsum('...ji,...jk ->...ik', win_mu, win_mu)    # :90
    inv = np.linalg.inv(win_var + (eps / win_size) * np.eye(d))   # :92
    X = np.einsum('...ij,...jk->...ik', winI - win_mu, inv)       # :94
    vals = np.eye(win_size) - (1.0 / win_size) * (1 + np.einsum('...ij,...kj->...ik', X, winI - win_mu))  # :95
    cols = np.tile(win_inds, win_size).ravel()              # :97
    rows = np.repeat(win_inds, win_size).ravel()            # :98
    return rows.astype(np.int64), cols.astype(np.int64), vals.ravel(), (h * w, h * w)   # :99-100


class V3Operator:
    """matting_v3.MattingLaplacian: COO build + sparse-dense matmul (matting_v3.py:27-51)."""

    def __init__(self, image, epsilon=1e-5, window_radius=1):
        image = np.asarray(image)
        self.size = image.shape
        self.rows, self.cols, self.vals, shp = v3_compute_laplacian(image, epsilon, window_radius)
        self.vals = self.vals.astype(image.dtype)           # :36-39 tf.cast(..., image.dtype)
        self.coo = scipy.sparse.coo_matrix((self.vals, (self.rows, self.cols)), shape=shp)
        self.csr = self.coo.tocsr()                         # duplicates summed: same product

    @property
    def shape(self):
        return self.coo.shape

    def matmul(self, x):
        return self.csr @ np.asarray(x)                     # :50-51
