"""Detector surface of the reference's src/detection package (ViTDetector / YOLOv8Detector), on the GPU."""

from . import ops  # noqa: F401  (registers the tensor-core building-block entry points)
from .vit_detector import DetrEngine, ViTDetector, input_shape, postprocess_tensors  # noqa: F401

__all__ = ["ViTDetector", "DetrEngine", "input_shape", "postprocess_tensors", "ops"]
