"""The two independent CPU restatements agree, the quirks of the reference are pinned as known
answers, and the committed golden vectors are what the oracle produces today.  CPU only."""
import numpy as np
import torch

import cases
from oracle import corr_np, raft_torch as rt, tf_ops, weights


def test_corr_numpy_vs_torch():
    f1, f2 = cases.fmaps(2, 8, 12, 64)
    a = corr_np.CorrBlock(f1, f2, 4, 4)
    b = rt.CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), 4, 4)
    for pa, pb in zip(a.corr_pyramid, b.corr_pyramid):
        np.testing.assert_allclose(pa, pb.numpy(), atol=1e-5, rtol=1e-5)
    for kind in ('grid', 'jitter', 'edge'):
        c = cases.lookup_coords(2, 8, 12, kind)
        ra = a.retrieve(c)
        # same pyramid on both sides -> the lookup itself must agree to rounding
        b.corr_pyramid = [torch.from_numpy(p) for p in a.corr_pyramid]
        rb = b.retrieve(torch.from_numpy(c)).numpy()
        np.testing.assert_allclose(ra, rb, atol=1e-6, rtol=1e-6)


def test_integer_coordinate_gives_exact_zero():
    """SURVEY.md trap 1 (corr.py:45-60): x or y integer => all four weights hold a zero factor."""
    img = np.arange(1, 21, dtype=np.float32).reshape(1, 4, 5, 1)
    coords = np.array([[[[2.0, 1.5], [1.5, 2.0], [2.0, 2.0], [1.5, 1.5]]]], dtype=np.float32)
    out = corr_np.bilinear_sampler(img, coords)[0, 0, :, 0]
    assert out[0] == 0 and out[1] == 0 and out[2] == 0
    assert out[3] == np.float32(0.25) * (img[0, 1, 1, 0] + img[0, 1, 2, 0] + img[0, 2, 1, 0] + img[0, 2, 2, 0])


def test_out_of_range_is_clamped_then_zero():
    """SURVEY.md trap 2 (corr.py:41-42): clamping lands on an integer coordinate => 0, never a gather OOB."""
    img = np.arange(1, 21, dtype=np.float32).reshape(1, 4, 5, 1)
    coords = np.array([[[[-3.2, 1.5], [7.9, 1.5], [1.5, -0.1], [1.5, 99.0], [4.0, 1.5]]]], dtype=np.float32)
    np.testing.assert_array_equal(corr_np.bilinear_sampler(img, coords)[0, 0, :, 0], np.zeros(5, np.float32))


def test_iteration0_level0_all_zero_and_tap_order():
    """Integer grid (iteration 0): every level-0 tap is 0 (trap 1).  Tap order (trap 3): channel a*9+b of a
    level samples at (x + a - r, y + b - r)."""
    f1, f2 = cases.fmaps(1, 8, 8, 32, seed=5)
    cb = corr_np.CorrBlock(f1, f2, 2, 4)
    out = cb.retrieve(cases.lookup_coords(1, 8, 8, 'grid'))
    assert np.all(out[..., :81] == 0)
    # non-integer query: compare one tap with a direct standard-bilinear evaluation
    c = np.full((1, 8, 8, 2), 3.25, np.float32)
    c[..., 1] = 4.5
    out = cb.retrieve(c)
    q = 2 * 8 + 5
    a, b = 6, 3                                    # x offset +2, y offset -1
    want = corr_np.standard_bilinear(cb.corr_pyramid[0][q:q + 1], np.array([[[[3.25 + 2, 4.5 - 1]]]], np.float32))
    np.testing.assert_allclose(out[0, 2, 5, a * 9 + b], want[0, 0, 0, 0], rtol=1e-6)


def test_pyramid_floors_odd_dims():
    """SURVEY.md trap 6 (corr.py:113): VALID pooling floors; 9x7 -> 4x3 -> 2x1."""
    f1, f2 = cases.fmaps(1, 9, 7, 16)
    cb = corr_np.CorrBlock(f1, f2, 3, 3)
    assert [p.shape[1:3] for p in cb.corr_pyramid] == [(9, 7), (4, 3), (2, 1)]
    np.testing.assert_allclose(cb.corr_pyramid[1][:, 0, 0, 0], cb.corr_pyramid[0][:, :2, :2, 0].mean(axis=(1, 2)),
                               rtol=1e-5, atol=1e-6)


def test_pooling_commutes_with_the_matmul():
    """Linearity used by the tensor-core path: pool(corr) == fmap1 . pool(fmap2)^T / sqrt(C)."""
    f1, f2 = cases.fmaps(1, 8, 12, 64)
    cb = corr_np.CorrBlock(f1, f2, 3, 4)
    f2l = f2
    for l in range(1, 3):
        f2l = tf_ops.avg_pool2d_2x2_valid(f2l)
        lin = corr_np.CorrBlock.correlation(f1, np.zeros_like(f1))  # shape helper unused
        n1 = f1.reshape(1, -1, 64)
        lvl = (n1 @ f2l.reshape(1, -1, 64).transpose(0, 2, 1)) / np.sqrt(np.float32(64))
        np.testing.assert_allclose(cb.corr_pyramid[l].reshape(lvl.shape), lvl, atol=2e-5, rtol=1e-5)
        del lin


def test_conv_same_padding_matches_literal():
    """Keras SAME incl. the asymmetric stride-2 case (trap 14): torch path vs literal NumPy conv."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 10, 12, 5)).astype(np.float32)
    for (kh, kw, s) in ((7, 7, 2), (3, 3, 2), (3, 3, 1), (1, 5, 1), (5, 1, 1), (1, 1, 2)):
        k = rng.standard_normal((kh, kw, 5, 4)).astype(np.float32)
        b = rng.standard_normal(4).astype(np.float32)
        pad = 'valid' if (kh, kw) == (1, 1) else 'same'
        lit = tf_ops.conv2d(x, k, b, strides=s, padding=pad)
        ops = rt.Ops({'c.kernel': k, 'c.bias': b})
        tor = ops.conv(torch.from_numpy(x).permute(0, 3, 1, 2), 'c', s, pad).permute(0, 2, 3, 1).numpy()
        np.testing.assert_allclose(tor, lit, atol=2e-5, rtol=1e-5)


def test_upflow8_is_half_pixel_bilinear():
    """trap 10 (corr.py:93-96)."""
    rng = np.random.default_rng(4)
    flow = rng.standard_normal((1, 3, 5, 2)).astype(np.float32)
    lit = corr_np.upflow8(flow)
    tor = rt.upflow8(torch.from_numpy(flow)).numpy()
    np.testing.assert_allclose(lit, tor, atol=1e-5, rtol=1e-5)


def test_goldens_are_current(golden):
    """The committed vectors are reproducible from the seeds (guards against silent oracle drift)."""
    f1, f2 = cases.fmaps(2, 8, 12, 64)
    cb = corr_np.CorrBlock(f1, f2, 4, 4)
    g = golden['corr_lookup']
    np.testing.assert_array_equal(cb.corr_pyramid[2], g['a_pyr2'])
    np.testing.assert_array_equal(cb.retrieve(cases.lookup_coords(2, 8, 12, 'edge')), g['a_lookup_edge'])
    p = weights.init_params('small', 1234, bias_scale=0.05, norm_jitter=0.1)
    im1, im2 = cases.images(1, 64, 128)
    preds = rt.forward(p, im1, im2, 'small', 3)
    np.testing.assert_allclose(np.stack([q.numpy() for q in preds]), golden['models']['small_64x128_it3'],
                               atol=2e-4, rtol=1e-5)


def test_fp32_is_far_inside_the_parity_gate_but_tf32_is_not():
    """DESIGN.md 'Precision': fp64 truth vs fp32 and vs TF32-rounded operands (small case)."""
    p = weights.init_params('raft', 1234)
    im1, im2 = cases.images(1, 64, 96)
    f64 = rt.forward(p, im1, im2, 'raft', 4, dtype=torch.float64)[-1]
    f32 = rt.forward(p, im1, im2, 'raft', 4)[-1].double()
    h = lambda t: t.half().float()
    sc = lambda t: (t * 256.0).half().float() / 256.0
    x2 = rt.forward(p, im1, im2, 'raft', 4, split=(h, h, sc, sc))[-1].double()
    t32 = rt.forward(p, im1, im2, 'raft', 4, quant=rt.tf32_trunc)[-1].double()
    assert (f32 - f64).abs().max() < 1e-4
    assert (x2 - f64).abs().max() < 2e-4          # fp16 hi/lo split: fp32-grade
    assert (t32 - f64).abs().max() > 1e-3         # plain TF32 operands break the <=1e-3 gate
