from . import sharpness, speckles, statistics
from .sharpness import sharpness_stack_stats, sharpness_stats
from .speckles import speckle_stack_stats, speckle_stats
from .statistics import distribution_moments

__all__ = ["sharpness", "sharpness_stats", "sharpness_stack_stats", "statistics", "speckles", "speckle_stats",
           "speckle_stack_stats", "distribution_moments"]
