"""NumPy restatements of the TensorFlow 2.3 ops the reference's hot path calls.

Test infrastructure (see oracle/__init__.py).  Each function states the TF op
and the reference call site it stands in for.  Semantics follow SURVEY.md
Appendix A.  Everything is float32 unless a dtype is passed in.
"""
import numpy as np

F32 = np.float32


def same_pads(n_in, k, s):
    """Keras/TF 'SAME' padding for one spatial dim -> (before, after, n_out).

    out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total//2.
    Asymmetric for even `in` with s=2 (extractor.py:26,95: 3x3 s2 -> 0/1, 7x7 s2 -> 2/3).
    """
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + k - n_in, 0)
    return total // 2, total - total // 2, n_out


def conv2d(x, kernel, bias=None, strides=1, padding='same'):
    """`tf.keras.layers.Conv2D` forward: NHWC input, HWIO kernel, cross-correlation.

    Call sites: update.py:10-11,22-24,43-49,73-76,91-95,138-140; extractor.py:26-27,37,95,102.
    Literal (slow) loop over kernel taps; used on small cases only.
    """
    x = np.asarray(x, F32)
    kh, kw, cin, cout = kernel.shape
    b, h, w, _ = x.shape
    if padding == 'same':
        pt, pb, ho = same_pads(h, kh, strides)
        pl, pr, wo = same_pads(w, kw, strides)
    else:
        pt = pb = pl = pr = 0
        ho = (h - kh) // strides + 1
        wo = (w - kw) // strides + 1
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((b, ho, wo, cout), F32)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (ho - 1) * strides + 1:strides, j:j + (wo - 1) * strides + 1:strides, :]
            out += np.tensordot(patch, kernel[i, j].astype(F32), axes=([3], [0])).astype(F32)
    if bias is not None:
        out = out + bias.astype(F32)
    return out.astype(F32)


def avg_pool2d_2x2_valid(x):
    """`tf.nn.avg_pool2d(x, 2, 2, 'VALID')` on NHWC (corr.py:113): floor on odd dims."""
    b, h, w, c = x.shape
    ho, wo = h // 2, w // 2
    x = x[:, :ho * 2, :wo * 2, :].reshape(b, ho, 2, wo, 2, c)
    return x.mean(axis=(2, 4), dtype=F32).astype(F32)


def gather_nd_batch1(params, idx):
    """`tf.gather_nd(params, idx, batch_dims=1)` with idx[..., (y, x)] (corr.py:63-66).

    params (M, H, W, C), idx (M, ..., 2) int32 -> (M, ..., C).
    """
    m = params.shape[0]
    lead = np.arange(m).reshape((m,) + (1,) * (idx.ndim - 2))
    return params[lead, idx[..., 0], idx[..., 1], :]


def extract_patches_3x3_same(x):
    """`tf.image.extract_patches(x, (1,3,3,1), (1,1,1,1), (1,1,1,1), 'SAME')` (model.py:55-59).

    Zero padding 1; depth index (ky*3 + kx)*C + c -- the order pinned by the
    reference's tests/test_model.py:20-21,40.
    """
    b, h, w, c = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    cols = [xp[:, ky:ky + h, kx:kx + w, :] for ky in range(3) for kx in range(3)]
    return np.concatenate(cols, axis=-1)


def extract_patches_valid(x, k):
    """VALID-padded k x k patches (only used to reproduce tests/test_model.py:28-33)."""
    b, h, w, c = x.shape
    ho, wo = h - k + 1, w - k + 1
    cols = [x[:, ky:ky + ho, kx:kx + wo, :] for ky in range(k) for kx in range(k)]
    return np.concatenate(cols, axis=-1)


def depth_to_space(x, bs):
    """`tf.nn.depth_to_space(x, bs)` NHWC (model.py:66): out[b, bs*y+by, bs*x+bx, c] = x[b,y,x,(by*bs+bx)*C+c]."""
    b, h, w, d = x.shape
    c = d // (bs * bs)
    x = x.reshape(b, h, w, bs, bs, c).transpose(0, 1, 3, 2, 4, 5)
    return x.reshape(b, h * bs, w * bs, c)


def softmax(x, axis):
    """`tf.nn.softmax` (model.py:52): max-subtracted."""
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x, dtype=F32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def resize_bilinear(x, new_h, new_w):
    """`tf.image.resize(x, size, 'bilinear')` TF2 semantics (corr.py:96): half-pixel centres,
    no antialias, source index clamped to [0, in-1]."""
    b, h, w, c = x.shape

    def axis_weights(n_in, n_out):
        scale = F32(n_in) / F32(n_out)
        src = (np.arange(n_out, dtype=F32) + F32(0.5)) * scale - F32(0.5)
        lo = np.floor(src)
        frac = (src - lo).astype(F32)
        lo_i = np.clip(lo.astype(np.int64), 0, n_in - 1)
        hi_i = np.clip(lo.astype(np.int64) + 1, 0, n_in - 1)
        return lo_i, hi_i, frac

    y0, y1, fy = axis_weights(h, new_h)
    x0, x1, fx = axis_weights(w, new_w)
    top = x[:, y0][:, :, x0] * (1 - fx)[None, None, :, None] + x[:, y0][:, :, x1] * fx[None, None, :, None]
    bot = x[:, y1][:, :, x0] * (1 - fx)[None, None, :, None] + x[:, y1][:, :, x1] * fx[None, None, :, None]
    return (top * (1 - fy)[None, :, None, None] + bot * fy[None, :, None, None]).astype(F32)
