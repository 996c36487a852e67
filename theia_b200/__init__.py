"""theia_b200 -- B200-native (sm_100a) implementation of the Theia distillation hot path.

Public surface mirrors the reference's `theia.models.rvfm.RobotVisionFM`."""
from .rvfm import RobotVisionFM, handle_feature_output  # noqa: F401

__all__ = ["RobotVisionFM", "handle_feature_output"]
