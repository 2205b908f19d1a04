"""Seeded synthetic inputs shared by the golden generator and the parity tests.

Inputs are never stored: both sides regenerate them from the same NumPy PCG64 seeds
(SURVEY.md section 8(d)).  Only oracle OUTPUTS are committed under tests/golden/.
"""
import numpy as np

F32 = np.float32


def images(b, h, w, seed0=0, seed1=1):
    """README.md:99-100 style inputs: uniform [0, 255) float32 image pairs."""
    im1 = np.random.default_rng(seed0).uniform(0, 255, (b, h, w, 3)).astype(F32)
    im2 = np.random.default_rng(seed1).uniform(0, 255, (b, h, w, 3)).astype(F32)
    return im1, im2


def fmaps(b, h, w, c, seed=10):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((b, h, w, c)).astype(F32), rng.standard_normal((b, h, w, c)).astype(F32)


def lookup_coords(b, h, w, kind, seed=2):
    """Query coordinates for CorrBlock.retrieve.

    'grid'    : the integer pixel grid (iteration 0: every level-0 tap is exactly 0)
    'jitter'  : grid + U(-8, 8)^2 -- in and out of range, non-integer
    'edge'    : a mix that pins the quirks: exact integers, half-integers, exact borders, far outside
    """
    gy, gx = np.meshgrid(np.arange(h, dtype=F32), np.arange(w, dtype=F32), indexing='ij')
    grid = np.tile(np.stack([gx, gy], axis=-1)[None], (b, 1, 1, 1)).astype(F32)
    if kind == 'grid':
        return grid
    rng = np.random.default_rng(seed)
    if kind == 'jitter':
        return (grid + rng.uniform(-8, 8, grid.shape)).astype(F32)
    if kind == 'edge':
        c = (grid + rng.uniform(-3, 3, grid.shape)).astype(F32)
        flat = c.reshape(-1, 2)
        n = flat.shape[0]
        flat[0::7] = np.round(flat[0::7])                       # exact integers
        flat[1::7] = np.round(flat[1::7]) + F32(0.5)            # half integers (level-1 integers)
        flat[2::7, 0] = F32(w - 1)                              # right border
        flat[3::7, 1] = F32(0)                                  # top border
        flat[4::7] = flat[4::7] + F32(100)                      # far outside
        flat[5::7] = -flat[5::7] - F32(50)
        flat[6 % n::11] = np.round(flat[6 % n::11] * 4) / 4     # quarter integers (level-2 integers)
        return flat.reshape(c.shape)
    raise ValueError(kind)


def update_inputs(variant, b, h, w, seed=20):
    hid, ctx, cch = (128, 128, 324) if variant == 'raft' else (96, 64, 196)
    rng = np.random.default_rng(seed)
    net = np.tanh(rng.standard_normal((b, h, w, hid))).astype(F32)
    inp = np.maximum(rng.standard_normal((b, h, w, ctx)), 0).astype(F32)
    corr = (rng.standard_normal((b, h, w, cch)) * 3).astype(F32)
    flow = (rng.standard_normal((b, h, w, 2)) * 4).astype(F32)
    return net, inp, corr, flow
