"""
Drop-in for ``barc4dip.signal`` on the B200 path: same public names, served by the CUDA library (``csrc/spectral.cu``,
``csrc/generic_dft.cuh``) through ``engine``. The 1-D helpers run the same kernels on (1, n) frames.
"""

from . import corr, fft, tracking

_EXPORTS = {
    fft: ("freq_axis1d", "freq_axes2d", "fft1d", "ifft1d", "psd1d", "fft2d", "ifft2d", "psd2d"),
    corr: ("xcorr1d", "autocorr1d", "xcorr2d", "autocorr2d"),
    tracking: ("track_translation", "phase_correlation", "template_matching"),
}
for _module, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)

__all__ = ["fft", "corr", "tracking"] + [n for names in _EXPORTS.values() for n in names]
del _module, _names, _name
