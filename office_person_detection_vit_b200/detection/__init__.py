"""Detector surface of the reference's src/detection package (ViTDetector / YOLOv8Detector), on the GPU."""

from . import ops  # noqa: F401  (registers the tensor-core building-block entry points)
