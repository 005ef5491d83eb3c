"""
Drop-in for ``barc4dip.signal`` on the B200 path: same public names, served by the CUDA library (``csrc/spectral.cu``,
``csrc/generic_dft.cuh``) through ``engine``. The 1-D helpers of the reference (fft1d, psd1d, xcorr1d, autocorr1d) are not
part of the stack-analysis hot path and are not provided.
"""

from . import corr, fft, tracking

_EXPORTS = {
    fft: ("freq_axes2d", "fft2d", "ifft2d", "psd2d"),
    corr: ("xcorr2d", "autocorr2d"),
    tracking: ("track_translation", "phase_correlation", "template_matching"),
}
for _module, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)

__all__ = ["fft", "corr", "tracking"] + [n for names in _EXPORTS.values() for n in names]
del _module, _names, _name
