from .data_models import AggregationResult, Detection, FrameResult

__all__ = ["AggregationResult", "Detection", "FrameResult"]
