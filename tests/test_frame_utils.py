"""Flow file formats (tf_raft/datasets/frame_utils.py:12-31, 102-120): byte-level known answers and round trips."""
import struct

import numpy as np
import pytest

from tf_raft_b200.datasets import read_flow, read_flow_kitti, write_flow, write_flow_kitti


def test_flo_known_bytes_and_round_trip(tmp_path):
    flow = np.arange(2 * 3 * 2, dtype=np.float32).reshape(2, 3, 2) - 2.5
    p = tmp_path / 'a.flo'
    write_flow(p, flow)
    raw = p.read_bytes()
    assert raw[:12] == struct.pack('<fii', 202021.25, 3, 2)                 # tag 'PIEH', width, height
    assert raw[12:20] == struct.pack('<ff', -2.5, -1.5)                     # (u, v) of pixel (0, 0), row-major
    np.testing.assert_array_equal(read_flow(p), flow)
    (tmp_path / 'bad.flo').write_bytes(struct.pack('<fii', 1.0, 3, 2) + raw[12:])
    with pytest.raises(ValueError):
        read_flow(tmp_path / 'bad.flo')


def test_kitti_png_round_trip_and_scaling(tmp_path):
    rng = np.random.default_rng(0)
    flow = (rng.integers(-200 * 64, 200 * 64, (5, 7, 2)) / 64.0).astype(np.float32)      # representable exactly
    valid = (rng.uniform(size=(5, 7)) > 0.3).astype(np.float32)
    p = tmp_path / 'k.png'
    write_flow_kitti(p, flow, valid)
    got, v = read_flow_kitti(p)
    np.testing.assert_array_equal(got, flow)
    np.testing.assert_array_equal(v, valid)
    try:                                             # cross-check against OpenCV's decoder when it is installed
        import cv2
        img = cv2.imread(str(p), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float32)
        np.testing.assert_array_equal((img[:, :, :2] - 2 ** 15) / 64.0, flow)
    except ImportError:
        pass
