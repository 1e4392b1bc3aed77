from .floormap_config import FloorMapConfig
from .homography import HomographyTransformer, TransformResult
from .piecewise_affine import PiecewiseAffineTransformer, PWATransformResult, ThinPlateSplineTransformer

__all__ = ["FloorMapConfig", "HomographyTransformer", "TransformResult", "PiecewiseAffineTransformer", "PWATransformResult",
           "ThinPlateSplineTransformer"]
