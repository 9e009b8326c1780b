"""B200-native scene-flow / motion-segmentation hot path (drop-in for moving_object_detector's
scene_flow_constructor construct() and scene_flow_clusterer clustering()), behind the C ABI of include/mod_b200.h."""
from ._lib import LIB_PATH, ModError, load  # noqa: F401
from .api import (CameraInfo, ClusterResult, ClustererConfig, Context, DisparityImage, MovingObject,  # noqa: F401
                  OBJECT_DTYPE, STATS_DTYPE, SceneFlowConstructorConfig, TransformMsg, objects_from_records)

__all__ = ["Context", "CameraInfo", "DisparityImage", "TransformMsg", "SceneFlowConstructorConfig", "ClustererConfig",
           "MovingObject", "ClusterResult", "ModError", "load", "LIB_PATH", "OBJECT_DTYPE", "STATS_DTYPE",
           "objects_from_records"]
