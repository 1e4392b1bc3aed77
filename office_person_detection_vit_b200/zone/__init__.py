from .zone_classifier import ZoneClassifier, ZoneTable

__all__ = ["ZoneClassifier", "ZoneTable"]
