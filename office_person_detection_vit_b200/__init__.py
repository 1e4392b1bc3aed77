"""B200-native (sm_100a) implementation of the Phase 2 -> 3 hot path of
Kizuna42/office-person-detection-vit: batched DETR-ResNet-50 person detection, foot-point
homography, polygon zone classification and per-timestamp zone counting.

The public classes keep the reference's engine surfaces:

    detection.ViTDetector                 <- src/detection (ViTDetector / YOLOv8Detector surface)
    transform.HomographyTransformer       <- src/transform/homography.py
    transform.FloorMapConfig              <- src/transform/floormap_config.py
    transform.PiecewiseAffineTransformer / ThinPlateSplineTransformer <- src/transform/piecewise_affine.py
    calibration.LensDistortionCorrector   <- src/calibration/lens_distortion.py (points only)
    zone.ZoneClassifier                   <- src/zone/zone_classifier.py
    aggregation.Aggregator                <- src/aggregation/aggregator.py
    models.Detection / FrameResult / ...  <- src/models/data_models.py

All arithmetic runs in hand-written CUDA behind the C ABI of include/opd_b200.h
(libopd_b200.so, loaded with ctypes).  There is no CPU fallback: a missing library or GPU raises.
"""

__version__ = "0.1.0"
