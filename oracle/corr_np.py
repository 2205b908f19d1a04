"""Literal NumPy restatement of tf_raft/layers/corr.py (reference @ 3c85f54).

Test infrastructure (see oracle/__init__.py).  Op-for-op: every TF call in the
reference has one NumPy line here, in the same order, so the quirks come out
the same way (SURVEY.md section 8 'Parity traps'):

  * floor/ceil corners => an integer (or clamped) coordinate gives weight 0 on
    all four corners => the sample is exactly 0              (corr.py:45-60)
  * clamp-then-sample, never zero-pad                        (corr.py:41-42)
  * tap (a, b) of the (2r+1)^2 window has x-offset d[a], y-offset d[b]
                                                             (corr.py:133-143)
  * divide by sqrt(C) after the fp32 matmul                  (corr.py:160-162)
"""
import numpy as np

from . import tf_ops

F32 = np.float32


def bilinear_sampler(image, coords):
    """corr.py:28-69.  image (M, H, W, 1); coords (M, P, Q, 2) xy-ordered -> (M, P, Q, 1)."""
    image = np.asarray(image, F32)
    coords = np.asarray(coords, F32)
    _, h, w, _ = image.shape
    gx, gy = coords[..., 0], coords[..., 1]                       # tf.unstack      :40
    gx = np.clip(gx, F32(0), F32(w - 1))                          # clip_by_value   :41
    gy = np.clip(gy, F32(0), F32(h - 1))                          #                 :42
    gx0, gx1 = np.floor(gx), np.ceil(gx)                          #                 :45-46
    gy0, gy1 = np.floor(gy), np.ceil(gy)                          #                 :47-48
    g00 = np.stack([gy0, gx0], axis=-1)                           #                 :51-54
    g01 = np.stack([gy0, gx1], axis=-1)
    g10 = np.stack([gy1, gx0], axis=-1)
    g11 = np.stack([gy1, gx1], axis=-1)
    c00 = ((gy1 - gy) * (gx1 - gx))[..., None]                    #                 :57-60
    c01 = ((gy1 - gy) * (gx - gx0))[..., None]
    c10 = ((gy - gy0) * (gx1 - gx))[..., None]
    c11 = ((gy - gy0) * (gx - gx0))[..., None]
    x00 = tf_ops.gather_nd_batch1(image, g00.astype(np.int32))    #                 :63-66
    x01 = tf_ops.gather_nd_batch1(image, g01.astype(np.int32))
    x10 = tf_ops.gather_nd_batch1(image, g10.astype(np.int32))
    x11 = tf_ops.gather_nd_batch1(image, g11.astype(np.int32))
    return (c00 * x00 + c01 * x01 + c10 * x10 + c11 * x11).astype(F32)   #          :68


def standard_bilinear(image, coords):
    """What `tfa.image.resampler` computes (corr.py:6-25, tests/layers/test_corr.py:23):
    ordinary bilinear interpolation with floor / floor+1 corners and zero outside.
    Used only to reproduce the reference's own sampler test."""
    image = np.asarray(image, F32)
    coords = np.asarray(coords, F32)
    m, h, w, _ = image.shape
    gx, gy = coords[..., 0], coords[..., 1]
    x0 = np.floor(gx)
    y0 = np.floor(gy)
    fx = gx - x0
    fy = gy - y0
    out = np.zeros(coords.shape[:-1] + (1,), F32)
    lead = np.arange(m).reshape((m,) + (1,) * (gx.ndim - 1))
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi = (x0 + dx).astype(np.int64)
            yi = (y0 + dy).astype(np.int64)
            ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            v = image[lead, np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1), 0]
            out[..., 0] += np.where(ok, v * wy * wx, 0).astype(F32)
    return out


def coords_grid(batch_size, height, width):
    """corr.py:72-90 -> (B, H, W, 2), last dim (x, y)."""
    gy, gx = np.meshgrid(np.arange(height, dtype=F32), np.arange(width, dtype=F32), indexing='ij')
    coords = np.stack([gx, gy], axis=-1)[None]
    return np.tile(coords, (batch_size, 1, 1, 1)).astype(F32)


def upflow8(flow):
    """corr.py:93-96: 8 * tf.image.resize(flow, (8h, 8w), 'bilinear')."""
    _, h, w, _ = flow.shape
    return (F32(8) * tf_ops.resize_bilinear(np.asarray(flow, F32), 8 * h, 8 * w)).astype(F32)


class CorrBlock:
    """corr.py:99-162."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.fmap1 = np.asarray(fmap1, F32)
        self.fmap2 = np.asarray(fmap2, F32)
        self.num_levels = num_levels
        self.radius = radius
        corr = self.correlation(self.fmap1, self.fmap2)                       # :106
        bs, h1, w1, _, h2, w2 = corr.shape
        corr = corr.reshape(bs * h1 * w1, h2, w2, 1)                          # :108
        self.corr_pyramid = [corr]                                            # :111
        for _ in range(num_levels - 1):
            corr = tf_ops.avg_pool2d_2x2_valid(corr)                          # :113
            self.corr_pyramid.append(corr)

    def retrieve(self, coords):
        r = self.radius
        coords = np.asarray(coords, F32)
        bs, h, w, _ = coords.shape
        out_pyramid = []
        for i in range(self.num_levels):
            corr = self.corr_pyramid[i]
            d = np.arange(-r, r + 1, dtype=F32)                               # :133
            dy, dx = np.meshgrid(d, d, indexing='ij')                         # :134
            delta = np.stack([dy, dx], axis=-1)                               # :136
            delta_lvl = delta.reshape(1, 2 * r + 1, 2 * r + 1, 2)             # :138
            centroid_lvl = coords.reshape(bs * h * w, 1, 1, 2) / F32(2 ** i)  # :141
            coords_lvl = (centroid_lvl + delta_lvl).astype(F32)               # :143
            s = bilinear_sampler(corr, coords_lvl)                            # :146
            out_pyramid.append(s.reshape(bs, h, w, -1))                       # :148
        return np.concatenate(out_pyramid, axis=-1)                           # :151

    @staticmethod
    def correlation(fmap1, fmap2):
        bs, h, w, nch = fmap1.shape
        f1 = fmap1.reshape(bs, h * w, nch)                                    # :156
        f2 = fmap2.reshape(bs, h * w, nch)                                    # :157
        corr = np.matmul(f1, f2.transpose(0, 2, 1)).astype(F32)               # :160
        corr = corr.reshape(bs, h, w, 1, h, w)                                # :161
        return (corr / np.sqrt(F32(nch))).astype(F32)                         # :162
