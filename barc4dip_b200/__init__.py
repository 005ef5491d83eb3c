"""
barc4dip_b200 -- B200-native (sm_100a CUDA) implementation of barc4dip's stack-analysis hot path.

Drop-in namespaces (same function signatures and array conventions as the reference):

    barc4dip_b200.signal         fft2d, psd2d, xcorr2d, autocorr2d, phase_correlation, track_translation
    barc4dip_b200.metrics        distribution_moments, tenengrad, laplacian_variance, amplitude, grain, bandwidth, ...
    barc4dip_b200.preprocessing  flat_field_correction

Stack-level, HBM-resident API: barc4dip_b200.engine, barc4dip_b200.stack, barc4dip_b200.parallel.
All arithmetic runs in libb4d.so (include/b4d.h); there is no CPU fallback.
"""

__version__ = "0.1.0"

from . import engine, metrics, preprocessing, signal  # noqa: E402,F401
from .metrics import distribution_moments  # noqa: E402,F401
