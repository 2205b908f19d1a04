"""Pin the oracle against every known answer the reference's own tests hold for this path
(SURVEY.md section 8(c)).  CPU only."""
import numpy as np
import torch

from oracle import corr_np, losses as olosses, raft_torch as rt, tf_ops


def test_sampler_matches_standard_bilinear_in_range():
    """reference tests/layers/test_corr.py:15-27: bilinear_sampler == tfa.image.resampler on in-range,
    non-integer coordinates, atol/rtol 1e-5 (image (4*32*32, 32, 32, 1) there; a slice of it here)."""
    rng = np.random.default_rng(1)
    m, h, w, r = 256, 32, 32, 4
    image = rng.standard_normal((m, h, w, 1)).astype(np.float32)
    cx = rng.uniform(0, w - 1, (m, 2 * r + 1, 2 * r + 1)).astype(np.float32)
    cy = rng.uniform(0, h - 1, (m, 2 * r + 1, 2 * r + 1)).astype(np.float32)
    coords = np.stack([cx, cy], axis=-1)
    valid = corr_np.standard_bilinear(image, coords)
    actual = corr_np.bilinear_sampler(image, coords)
    np.testing.assert_allclose(actual, valid, atol=1e-5, rtol=1e-5)
    actual_t = rt.bilinear_sampler(torch.from_numpy(image), torch.from_numpy(coords)).numpy()
    np.testing.assert_allclose(actual_t, valid, atol=1e-5, rtol=1e-5)


def test_extract_patches_and_depth_to_space_known_answer():
    """reference tests/test_model.py:14-41."""
    flow = np.array([[[1, 2], [3, 4], [5, 6]], [[7, 8], [9, 10], [11, 12]], [[13, 14], [15, 16], [17, 18]]],
                    dtype=np.float32)
    flow_unfold = np.array([[[1, 2, 3, 4, 7, 8, 9, 10], [3, 4, 5, 6, 9, 10, 11, 12]],
                            [[7, 8, 9, 10, 13, 14, 15, 16], [9, 10, 11, 12, 15, 16, 17, 18]]], dtype=np.float32)
    up_flow = np.array([[[1, 2], [3, 4], [3, 4], [5, 6]], [[7, 8], [9, 10], [9, 10], [11, 12]],
                        [[7, 8], [9, 10], [9, 10], [11, 12]], [[13, 14], [15, 16], [15, 16], [17, 18]]],
                       dtype=np.float32)
    unfold = tf_ops.extract_patches_valid(flow[None], 2)
    np.testing.assert_allclose(unfold[0], flow_unfold)
    np.testing.assert_allclose(tf_ops.depth_to_space(unfold, 2)[0], up_flow)


def test_upsample_flow_orders_follow_the_pinned_ops():
    """RAFT.upsample_flow (model.py:39-66) restated with the literal ops equals the torch restatement."""
    rng = np.random.default_rng(3)
    b, h, w = 2, 3, 4
    flow = rng.standard_normal((b, h, w, 2)).astype(np.float32)
    mask = rng.standard_normal((b, h, w, 576)).astype(np.float32)
    m = tf_ops.softmax(mask.reshape(b, h, w, 8, 8, 9, 1), axis=5)
    up = tf_ops.extract_patches_3x3_same(8 * flow).reshape(b, h, w, 1, 1, 9, 2)
    lit = tf_ops.depth_to_space((m * up).sum(axis=5).reshape(b, h, w, -1), 8)
    tor = rt.upsample_flow(torch.from_numpy(flow), torch.from_numpy(mask)).numpy()
    np.testing.assert_allclose(tor, lit, atol=2e-6, rtol=1e-6)


def _loss_fixture():
    flow_gt = np.array([[[0, 1], [0, 2], [0, 3]], [[0, 4], [0, 5], [0, 6]], [[0, 7], [0, 8], [0, 9]]]) - 0.1
    valid = np.array([[True, True, True], [True, True, True], [True, True, False]])
    flow_gt, valid = flow_gt[None], valid[None]
    return (flow_gt, valid), [np.zeros_like(flow_gt) for _ in range(6)]


def test_sequence_loss_known_answer():
    """reference tests/losses/test_losses.py:27-45."""
    (flow_gt, valid), preds = _loss_fixture()
    expect = sum(0.8 ** (6 - i - 1) * np.mean(valid[..., None] * np.abs(p - flow_gt)) for i, p in enumerate(preds))
    np.testing.assert_almost_equal(olosses.sequence_loss((flow_gt, valid), preds, gamma=0.8), expect)


def test_end_point_error_known_answer():
    """reference tests/losses/test_losses.py:48-67."""
    (flow_gt, valid), preds = _loss_fixture()
    info = olosses.end_point_error((flow_gt, valid), preds[-1])
    np.testing.assert_almost_equal(info['epe'], np.mean(np.arange(1, 9) - 0.1), decimal=2)
    for k, v in (('u1', 1 / 8), ('u3', 3 / 8), ('u5', 5 / 8)):
        np.testing.assert_almost_equal(info[k], v, decimal=2)


def test_model_output_contract():
    """reference tests/test_model.py:44-77: list length = iters, each (B, H, W, 2) -- at the reference's
    64x96 image size (batch 1 and fewer iterations to keep the CPU suite short)."""
    from oracle import weights
    import cases
    for variant in ('raft', 'small'):
        p = weights.init_params(variant, 1)
        im1, im2 = cases.images(1, 64, 96)
        preds = rt.forward(p, im1, im2, variant, iters=2)
        assert len(preds) == 2
        for f in preds:
            assert tuple(f.shape) == (1, 64, 96, 2)


def test_end_point_error_metric_class_known_answer():
    """reference tests/losses/test_losses.py:70-95 on the product's EndPointError metric (losses.py:46-85)."""
    import torch
    from tf_raft_b200.losses import EndPointError
    flow_gt = (np.array([[[0, 1], [0, 2], [0, 3]], [[0, 4], [0, 5], [0, 6]], [[0, 7], [0, 8], [0, 9]]]) - 0.1)[None].astype(np.float32)
    valid = np.array([[True, True, True], [True, True, True], [True, True, False]])[None]
    preds = [torch.zeros(1, 3, 3, 2) for _ in range(6)]
    m = EndPointError()
    for _ in range(3):
        m.update_state([torch.from_numpy(flow_gt), torch.from_numpy(valid)], preds)
    info = m.result()
    # valid pixels have gt (−0.1, k − 0.1), k = 1..8: epe = sqrt(0.01 + (k − 0.1)^2)
    k = np.arange(1, 9) - 0.1
    epe = np.sqrt(0.01 + k ** 2)
    np.testing.assert_almost_equal(info['epe'], epe.mean(), decimal=5)
    np.testing.assert_almost_equal(info['u1'], (epe < 1).mean(), decimal=6)
    np.testing.assert_almost_equal(info['u3'], (epe < 3).mean(), decimal=6)
    np.testing.assert_almost_equal(info['u5'], (epe < 5).mean(), decimal=6)
    assert m.count == 3
