"""``import pointconv_util`` (models_bid_pointconv.py:7-9, loss_functions.py:1) -> the kdpc layers."""
from kd_pointcloud_b200.pointconv_util import *  # noqa: F401,F403
from kd_pointcloud_b200.pointconv_util import LEAKY_RATE, use_bn, pointnet2_utils  # noqa: F401
