"""Drop-in mirrors of the reference's ``analyzers`` package for the technical metrics."""
from .image_cache import ImageCache
from .technical import TechnicalAnalyzer

__all__ = ["ImageCache", "TechnicalAnalyzer"]
