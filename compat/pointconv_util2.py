"""``import pointconv_util2`` (models_bid_lighttoken_res.py:6-7): the hot classes of the reference's
pointconv_util2.py are textually identical to pointconv_util.py."""
from kd_pointcloud_b200.pointconv_util import *  # noqa: F401,F403
from kd_pointcloud_b200.pointconv_util import LEAKY_RATE, use_bn, pointnet2_utils  # noqa: F401
