"""Correlation volume, pyramid and lookup -- host-side mirror of tf_raft/layers/corr.py.

Same names, argument meaning and error behaviour as the reference module, with torch CUDA
tensors in place of TF tensors.  Every function hands device pointers to libraft_b200.so
(include/raft_b200.h); nothing here computes on the host.
"""
import ctypes

import torch
import torch.nn.functional as F

from .. import _lib


def tfa_sampler(image, coords, mask=False):
    """Reference corr.py:6-25: `tfa.image.resampler` = ordinary bilinear interpolation with zero
    outside.  Not on the model's path (the reference never calls it); kept as the API alias."""
    if mask:
        raise NotImplementedError("mask is not implemented for True")
    m, h, w, _ = image.shape
    gx = 2 * coords[..., 0] / max(w - 1, 1) - 1
    gy = 2 * coords[..., 1] / max(h - 1, 1) - 1
    grid = torch.stack([gx, gy], dim=-1)
    out = F.grid_sample(image.permute(0, 3, 1, 2), grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    return out.permute(0, 2, 3, 1)


def bilinear_sampler(image, coords):
    """Reference corr.py:28-69.  image (M, H, W, 1), coords (M, P, Q, 2) xy -> (M, P, Q, 1).

    floor/ceil corners: a sample whose clamped x or y is an integer is exactly 0."""
    image = _lib.f32c(image)
    coords = _lib.f32c(coords)
    m, h, w, c = image.shape
    if c != 1:
        raise ValueError('bilinear_sampler expects a single-channel image (M, H, W, 1)')
    p = coords.shape[1] * coords.shape[2]
    out = torch.empty(coords.shape[:-1] + (1,), dtype=torch.float32, device=image.device)
    _lib.check(_lib.lib().raft_b200_bilinear_sampler(_lib.ptr(image), _lib.ptr(coords), m, h, w, p, _lib.ptr(out),
                                                     _lib.stream()), 'bilinear_sampler')
    return out


def coords_grid(batch_size, height, width, device=None):
    """Reference corr.py:72-90 -> (B, H, W, 2) with [..., 0] = x, [..., 1] = y."""
    device = torch.device('cuda') if device is None else torch.device(device)
    out = torch.empty((batch_size, height, width, 2), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().raft_b200_coords_grid(batch_size, height, width, _lib.ptr(out), _lib.stream()),
                   'coords_grid')
    return out


def upflow8(flow, mode='bilinear'):
    """Reference corr.py:93-96: 8 * tf.image.resize(flow, (8h, 8w), 'bilinear') (half-pixel centres)."""
    if mode != 'bilinear':
        raise NotImplementedError("only mode='bilinear' is implemented")
    flow = _lib.f32c(flow)
    b, h, w, _ = flow.shape
    out = torch.empty((b, 8 * h, 8 * w, 2), dtype=torch.float32, device=flow.device)
    _lib.check(_lib.lib().raft_b200_upflow8(_lib.ptr(flow), b, h, w, _lib.ptr(out), _lib.stream()), 'upflow8')
    return out


class CorrBlock:
    """Reference corr.py:99-162.  Plain class; the constructor builds the whole pyramid.

    Attributes as in the reference: fmap1, fmap2, num_levels, radius, corr_pyramid (list of
    `num_levels` tensors (B*h*w, h>>l, w>>l, 1)).  `precision` ('f16x2' tcgen05 path, or 'fp32'
    CUDA-core path) is a keyword-only extra.
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, *, precision=None):
        self.fmap1 = fmap1
        self.fmap2 = fmap2
        self.num_levels = num_levels
        self.radius = radius
        self.precision = _lib.resolve_precision(precision)
        f1, f2 = _lib.f32c(fmap1), _lib.f32c(fmap2)
        if f1.shape != f2.shape or f1.dim() != 4:
            raise ValueError(f'fmap1/fmap2 must both be (B, h, w, C); got {tuple(f1.shape)} and {tuple(f2.shape)}')
        b, h, w, c = f1.shape
        L = _lib.lib()
        sizes = (ctypes.c_size_t * num_levels)()
        _lib.check(L.raft_b200_corr_pyramid_sizes(b, h, w, num_levels, sizes), 'corr_pyramid_sizes')
        self.corr_pyramid = [torch.empty((b * h * w, h >> l, w >> l, 1), dtype=torch.float32, device=f1.device)
                             for l in range(num_levels)]
        nbytes = ctypes.c_size_t()
        _lib.check(L.raft_b200_corr_workspace_bytes(b, h, w, c, num_levels, self.precision, ctypes.byref(nbytes)),
                   'corr_workspace_bytes')
        ws = _lib.workspace(nbytes.value, f1.device)
        with torch.cuda.device(f1.device):
            _lib.check(L.raft_b200_corr_pyramid_build(_lib.ptr(f1), _lib.ptr(f2), b, h, w, c, num_levels,
                                                      _lib.ptr_array(self.corr_pyramid), _lib.ptr(ws), ws.numel(),
                                                      self.precision, _lib.stream()), 'corr_pyramid_build')
        self._ws = ws          # keep alive until the stream has consumed it
        self._shape = (b, h, w)

    def retrieve(self, coords):
        """coords (B, h, w, 2) xy -> (B, h, w, num_levels*(2r+1)^2)."""
        coords = _lib.f32c(coords)
        b, h, w, _ = coords.shape
        if (b, h, w) != self._shape:
            raise ValueError(f'coords shape {tuple(coords.shape)} does not match the correlation volume {self._shape}')
        nch = self.num_levels * (2 * self.radius + 1) ** 2
        out = torch.empty((b, h, w, nch), dtype=torch.float32, device=coords.device)
        with torch.cuda.device(coords.device):
            _lib.check(_lib.lib().raft_b200_corr_lookup(_lib.ptr_array(self.corr_pyramid), _lib.ptr(coords), b, h, w,
                                                        self.num_levels, self.radius, _lib.ptr(out), nch,
                                                        _lib.stream()), 'corr_lookup')
        return out

    def correlation(self, fmap1, fmap2):
        """Reference corr.py:154-162 -> (B, h, w, 1, h, w) = fmap1 . fmap2^T / sqrt(C)."""
        block = CorrBlock(fmap1, fmap2, num_levels=1, radius=self.radius, precision=self.precision)
        b, h, w = block._shape
        return block.corr_pyramid[0].reshape(b, h, w, 1, h, w)
