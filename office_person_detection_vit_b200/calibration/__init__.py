from .lens_distortion import CameraIntrinsics, DistortionParams, LensDistortionCorrector

__all__ = ["CameraIntrinsics", "DistortionParams", "LensDistortionCorrector"]
