"""
Drop-in for ``barc4dip.metrics``: per-frame metric functions and the four aggregators (single frame / stack, speckle /
sharpness), evaluated by batched kernels on HBM-resident stacks (``engine``, ``stack``).
"""

from . import common, sharpness, speckles, statistics

distribution_moments = statistics.distribution_moments
speckle_stats, speckle_stack_stats = speckles.speckle_stats, speckles.speckle_stack_stats
sharpness_stats, sharpness_stack_stats = sharpness.sharpness_stats, sharpness.sharpness_stack_stats

__all__ = ["statistics", "speckles", "sharpness", "common", "distribution_moments", "speckle_stats", "speckle_stack_stats",
           "sharpness_stats", "sharpness_stack_stats"]
