"""Crop-or-pad (reference datasets/dataset.py:323-334, training.py:73-81): TF's documented centring rule, checked on
hand-computed cases and by round trips (pure indexing, runs on CPU tensors)."""
import numpy as np
import pytest
import torch

import tf_raft_b200 as T


def test_sintel_frame_pads_to_448_and_crops_back():
    g = torch.Generator().manual_seed(0)
    img = torch.rand((2, 436, 1024, 3), generator=g)
    padded = T.resize_with_crop_or_pad(img, 448, 1024)
    assert padded.shape == (2, 448, 1024, 3)
    assert torch.equal(padded[:, 6:442], img)                 # (448-436)//2 = 6 rows before, 6 after
    assert padded[:, :6].abs().max() == 0 and padded[:, 442:].abs().max() == 0
    assert torch.equal(T.resize_with_crop_or_pad(padded, 436, 1024), img)


@pytest.mark.parametrize('size,target,crop0,pad0', [
    (5, 8, 0, 1),    # diff 3: pad 1 before, 2 after
    (8, 5, 1, 0),    # diff -3: -(-3)//2 = 1 cropped before, 2 after
    (7, 8, 0, 0),    # diff 1: pad 0 before, 1 after
    (8, 7, 0, 0),    # diff -1: 1//2 = 0 cropped before, 1 after
    (6, 6, 0, 0),
])
def test_offsets_follow_tf_floor_division(size, target, crop0, pad0):
    x = torch.arange(size, dtype=torch.float32).reshape(1, size, 1) + 1.0      # (H=1, W=size, C=1), values 1..size
    y = T.resize_with_crop_or_pad(x, 1, target).reshape(-1)
    want = np.zeros(target, dtype=np.float32)
    keep = min(size, target)
    want[pad0:pad0 + keep] = np.arange(size, dtype=np.float32)[crop0:crop0 + keep] + 1.0
    assert np.array_equal(y.numpy(), want)


def test_crop_or_padder_pads_valid_with_zeros_and_mixed_axes():
    f = T.CropOrPadder((8, 4))
    im1 = torch.ones((6, 6, 3)); im2 = 2 * torch.ones((6, 6, 3))
    flow = 3 * torch.ones((6, 6, 2)); valid = torch.ones((6, 6))
    a, b, fl, v = f(im1, im2, flow, valid)
    assert a.shape == (8, 4, 3) and b.shape == (8, 4, 3) and fl.shape == (8, 4, 2) and v.shape == (8, 4)
    assert v[0].sum() == 0 and v[7].sum() == 0 and v[1:7].min() == 1         # height padded (1 before, 1 after), width cropped
    assert fl[1:7].min() == 3 and b[1:7].min() == 2


def test_pad_to_multiple_round_trip_and_errors():
    x = torch.rand((1, 37, 50, 2))
    p, (h, w) = T.pad_to_multiple(x, 8)
    assert p.shape == (1, 40, 56, 2) and (h, w) == (37, 50)
    assert torch.equal(T.resize_with_crop_or_pad(p, h, w), x)
    q, _ = T.pad_to_multiple(torch.rand((16, 24, 3)))
    assert q.shape == (16, 24, 3)
    with pytest.raises(ValueError):
        T.resize_with_crop_or_pad(torch.rand(4, 4), 8, 8)
    with pytest.raises(ValueError):
        T.resize_with_crop_or_pad(torch.rand(4, 4, 1), 0, 8)
