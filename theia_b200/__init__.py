"""theia_b200 -- B200-native (sm_100a) implementation of the Theia distillation hot path.

Public surface mirrors the reference: `RobotVisionFM` (theia.models.rvfm), `theia_b200.optim.FlatAdamW` (the fused
optimizer tail), `theia_b200.data.ingest_targets` (target transpose + z-score) and `theia_b200.teachers` (the
`get_*_feature` wrappers of theia.foundation_models for the DINOv2 / CLIP / ViT teachers).  Everything computes
through `libtheia_b200.so` (include/theia_b200.h); there is no PyTorch fallback."""
from .rvfm import RobotVisionFM, handle_feature_output  # noqa: F401

__all__ = ["RobotVisionFM", "handle_feature_output"]
