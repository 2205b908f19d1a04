"""Centred crop-or-pad to a multiple-of-8 frame size -- the only preprocessing the hot path depends on.

The model needs H and W divisible by 8 (trap 7: `initialize_flow` uses `h // 8` while the
SAME-padded stride-2 encoder yields `ceil(h / 8)`), so the reference pads Sintel's 436x1024 frames to 448x1024 with
`tf.image.resize_with_crop_or_pad` before the forward and crops the predicted flow back
(tf_raft/datasets/dataset.py:323-334, tf_raft/training.py:73-81).  These helpers restate that op for torch tensors
(any device; pure indexing, no arithmetic), so BASELINE config 3 runs through the same API.
"""
import torch


def _offsets(size, target):
    """(crop offset, pad offset) of tf.image.resize_with_crop_or_pad along one axis: `max(-diff // 2, 0)` and
    `max(diff // 2, 0)` with Python floor division, diff = target - size."""
    diff = target - size
    return max(-diff // 2, 0), max(diff // 2, 0)


def resize_with_crop_or_pad(x, target_height, target_width):
    """`tf.image.resize_with_crop_or_pad(x, target_height, target_width)` for `(..., H, W, C)` tensors.

    Central crop where the input is larger, central zero padding where it is smaller (the odd pixel of an odd
    difference goes after / is cropped after, as in TF)."""
    if x.dim() < 3:
        raise ValueError('expected a (..., H, W, C) tensor')
    if target_height <= 0 or target_width <= 0:
        raise ValueError('target_height and target_width must be positive')
    h, w = x.shape[-3], x.shape[-2]
    ch, ph = _offsets(h, target_height)
    cw, pw = _offsets(w, target_width)
    keep_h, keep_w = min(h, target_height), min(w, target_width)
    cropped = x[..., ch:ch + keep_h, cw:cw + keep_w, :]
    if keep_h == target_height and keep_w == target_width:
        return cropped.contiguous()
    out = torch.zeros(x.shape[:-3] + (target_height, target_width, x.shape[-1]), dtype=x.dtype, device=x.device)
    out[..., ph:ph + keep_h, pw:pw + keep_w, :] = cropped
    return out


def CropOrPadder(target_size):
    """Reference datasets/dataset.py:323-334: returns `f(image1, image2, flow, valid)` that crop-or-pads all four
    (valid is `(..., H, W)` and is padded with zeros, i.e. padded pixels are invalid)."""
    th, tw = target_size

    def f(image1, image2, flow, valid):
        image1 = resize_with_crop_or_pad(image1, th, tw)
        image2 = resize_with_crop_or_pad(image2, th, tw)
        flow = resize_with_crop_or_pad(flow, th, tw)
        valid = resize_with_crop_or_pad(valid.unsqueeze(-1), th, tw).squeeze(-1)
        return image1, image2, flow, valid
    return f


def pad_to_multiple(x, multiple=8):
    """Smallest centred zero padding of `(..., H, W, C)` that makes H and W multiples of `multiple`; returns the
    padded tensor and the original `(H, W)` for `resize_with_crop_or_pad(flow, H, W)` afterwards."""
    h, w = x.shape[-3], x.shape[-2]
    th, tw = -(-h // multiple) * multiple, -(-w // multiple) * multiple
    return resize_with_crop_or_pad(x, th, tw), (h, w)
