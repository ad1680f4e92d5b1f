"""``from pointnet2 import pointnet2_utils`` (pointconv_util.py:14) -> the kdpc implementation."""
from kd_pointcloud_b200.pointnet2_utils import *  # noqa: F401,F403
from kd_pointcloud_b200.pointnet2_utils import (BallQuery, FurthestPointSampling, GatherOperation, GroupAll,  # noqa: F401
                                                GroupingOperation, QueryAndGroup, ThreeInterpolate, ThreeNN,
                                                ball_query, furthest_point_sample, gather_operation,
                                                grouping_operation, three_interpolate, three_nn)
