from . import corr, fft, tracking
from .corr import autocorr2d, xcorr2d
from .fft import fft2d, freq_axes2d, ifft2d, psd2d
from .tracking import phase_correlation, template_matching, track_translation

__all__ = ["fft", "corr", "tracking", "freq_axes2d", "fft2d", "ifft2d", "psd2d", "xcorr2d", "autocorr2d",
           "phase_correlation", "template_matching", "track_translation"]
