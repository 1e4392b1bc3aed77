from .floormap_config import FloorMapConfig
from .homography import HomographyTransformer, TransformResult

__all__ = ["FloorMapConfig", "HomographyTransformer", "TransformResult"]
