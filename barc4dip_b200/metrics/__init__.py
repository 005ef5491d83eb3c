from . import statistics
from .statistics import distribution_moments

__all__ = ["statistics", "distribution_moments"]
