"""Drop-in for the reference's evaluation_utils.py (put compat/ first on sys.path): same names, GPU evaluation."""
from kd_pointcloud_b200.evaluation_utils import evaluate_2d, evaluate_3d, scene_flow_metrics, MetricMeter  # noqa: F401
