"""File formats of the reference's data pipeline (tf_raft/datasets/frame_utils.py): Middlebury .flo and KITTI flow PNGs."""
from .frame_utils import read_flow, read_flow_kitti, write_flow, write_flow_kitti  # noqa: F401
