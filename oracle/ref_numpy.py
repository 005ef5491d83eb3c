"""
CPU oracle for the barc4dip stack-analysis hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the reference algorithms that the CUDA
path in ``barc4dip_b200`` replaces.  It exists to CHECK the product, never to be
the product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``barc4dip_b200/`` imports ``oracle``.

Parity pinning: the reference ships no tests or golden vectors, so
every function here is pinned against outputs of the *reference itself*, executed
in the build container by ``oracle/make_golden.py`` and committed under
``tests/golden/`` (checked by ``tests/test_oracle_golden.py``).  The single
exception is ``temporal_moments``: the reference has no such function, so that
row is "parity unpinned" -- it lifts the definitions of ``distribution_moments``
along axis 0.

The arithmetic the reference delegates to third-party libraries (numpy.fft,
scipy.ndimage.sobel/laplace, scipy.stats.describe, np.nanpercentile, np.median,
scipy's linear RegularGridInterpolator; all unpinned in the reference's
pyproject.toml:20-29) is called through the same libraries here, so the CPU
baseline timed from this file costs what the reference costs.  Explicit
restatements of those primitives (``sobel_reflect``, ``laplace_reflect``,
``bilinear_polar_mean``, ``percentile_linear``) document what the kernels compute
and are cross-checked against the library calls in the CPU test-suite.

"ref:" comments give file:line under /root/reference/src/barc4dip/.
"""

from __future__ import annotations

import math

import numpy as np
from scipy import ndimage
from scipy.interpolate import RegularGridInterpolator
from scipy.stats import describe

INV_E = 1.0 / math.e


# --------------------------------------------------------------------------------------
# signal.fft
# --------------------------------------------------------------------------------------

def freq_axes2d(shape, dx=1.0, dy=1.0):
    """Shifted frequency axes (fx, fy).  ref: signal/fft.py:58-96."""
    ny, nx = shape
    return (np.fft.fftshift(np.fft.fftfreq(int(nx), d=dx)),
            np.fft.fftshift(np.fft.fftfreq(int(ny), d=dy)))


def fft2d(image, dx=1.0, dy=1.0):
    """DC-centred unnormalised forward DFT.  ref: signal/fft.py:198-237 (fftshift(fft2) at :236)."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    fx, fy = freq_axes2d(img.shape, dx, dy)
    return np.fft.fftshift(np.fft.fft2(img)), fx, fy


def psd2d(image, dx=1.0, dy=1.0, scale=True):
    """|F|^2, optionally times dx*dy/(nx*ny); no mean removal, no window.  ref: signal/fft.py:261-309."""
    F, fx, fy = fft2d(image, dx, dy)
    P = np.abs(F) ** 2
    if scale:
        ny, nx = P.shape
        P = P * ((dx * dy) / (float(nx) * float(ny)))
    return P, fx, fy


# --------------------------------------------------------------------------------------
# signal.corr
# --------------------------------------------------------------------------------------

def lag_axis(n, step=1.0):
    """(arange(n) - n//2) * step.  ref: signal/common.py:89-90."""
    return (np.arange(n, dtype=float) - (n // 2)) * float(step)


def xcorr2d(a, b, dx=1.0, dy=1.0, remove_mean=True, standardize=False, normalize="peak"):
    """Circular cross-correlation via FFT, zero lag centred.  ref: signal/corr.py:169-253.

    float64 throughout (:213-214); means removed (:225-227); optional /std (:229-235);
    fftshift(ifft2(Fa * conj(Fb))) (:237-241); real_if_close(tol=1000) (:242);
    "peak" divides by max|corr| when positive (:247-251).
    """
    aa = np.asarray(a, dtype=float)
    bb = np.asarray(b, dtype=float)
    if aa.ndim != 2 or bb.ndim != 2:
        raise ValueError("a and b must be 2D arrays.")
    if aa.shape != bb.shape:
        raise ValueError("a and b must have the same shape.")
    ny, nx = aa.shape
    if remove_mean:
        aa = aa - float(aa.mean())
        bb = bb - float(bb.mean())
    if standardize:
        sa, sb = float(aa.std()), float(bb.std())
        aa = aa / sa if sa > 0 else aa
        bb = bb / sb if sb > 0 else bb
    c = np.fft.fftshift(np.fft.ifft2(np.fft.fft2(aa) * np.conj(np.fft.fft2(bb))))
    c = np.real_if_close(c, tol=1000)
    if normalize == "peak":
        m = float(np.max(np.abs(c)))
        if m > 0:
            c = c / m
    elif normalize != "none":
        raise ValueError(f"Invalid normalize='{normalize}'. Use 'none' or 'peak'.")
    return c, lag_axis(nx, dx), lag_axis(ny, dy)


def autocorr2d(a, dx=1.0, dy=1.0, remove_mean=True, standardize=False, normalize="peak"):
    """xcorr2d(a, a) forced real (raises on a significant imaginary part).  ref: signal/corr.py:256-320."""
    ac, xl, yl = xcorr2d(a, a, dx, dy, remove_mean, standardize, normalize)
    if np.iscomplexobj(ac):
        im, re = float(np.max(np.abs(ac.imag))), float(np.max(np.abs(ac.real)))
        if im > 1e-10 * max(re, 1.0):
            raise ValueError("autocorr2d returned significant imaginary part")
        ac = ac.real
    return ac, xl, yl


# --------------------------------------------------------------------------------------
# signal.tracking  (phase correlation, internal backend)
# --------------------------------------------------------------------------------------

def centered_roi_slices(image_shape, size_yx):
    """Centred odd-sized ROI; even sizes raise.  ref: geometry/roi.py:44-106 (odd check :81-82)."""
    H, W = image_shape
    sy, sx = size_yx
    if sy <= 0 or sx <= 0:
        raise ValueError("ROI sizes must be positive.")
    if sy % 2 == 0 or sx % 2 == 0:
        raise ValueError("ROI sizes must be odd for symmetry.")
    y0, x0 = H // 2 - sy // 2, W // 2 - sx // 2
    y1, x1 = y0 + sy, x0 + sx
    if y0 < 0 or y1 > H or x0 < 0 or x1 > W:
        raise ValueError("ROI exceeds image bounds.")
    return slice(y0, y1), slice(x0, x1)


def _float2d(a, name):
    """ints -> float32, floats keep their precision.  ref: signal/tracking.py:299-305."""
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError(f"{name} must be a 2D array.")
    return a if np.issubdtype(a.dtype, np.floating) else a.astype(np.float32)


def zscore2d(a, eps):
    """(a - nanmean) / (nanstd + eps); eps is added to the std.  ref: signal/tracking.py:308-311."""
    return (a - float(np.nanmean(a))) / (float(np.nanstd(a)) + eps)


def peak_subpixel_taylor(c, i, j):
    """3x3 quadratic refinement exactly as the reference writes it.  ref: signal/tracking.py:324-375.

    NOTE the returned pair: the first value is the *x* Newton step and the second the *y*
    step, yet the caller adds the first to dy and the second to dx (:293-295).  The swap is
    reference behaviour and is reproduced on purpose.
    Arithmetic stays in the map's dtype (float32 maps -> float32 scalars under NEP 50).
    """
    ny, nx = c.shape
    if i <= 0 or i >= ny - 1 or j <= 0 or j >= nx - 1:
        return 0.0, 0.0
    gy = (c[i + 1, j] - c[i - 1, j]) / 2.0
    gyy = c[i + 1, j] + c[i - 1, j] - 2.0 * c[i, j]
    gx = (c[i, j + 1] - c[i, j - 1]) / 2.0
    gxx = c[i, j + 1] + c[i, j - 1] - 2.0 * c[i, j]
    gxy = (c[i + 1, j + 1] - c[i + 1, j - 1] - c[i - 1, j + 1] + c[i - 1, j - 1]) / 4.0
    det = gxx * gyy - gxy * gxy
    if det == 0.0:
        return 0.0, 0.0
    inv = 1.0 / det
    return float(-(gyy * gx - gxy * gy) * inv), float(-(gxx * gy - gxy * gx) * inv)


def phase_correlation(template, image, slices_yx=None, subpixel=True, eps=1e-9):
    """FFT phase correlation of a template ROI against a full frame.  ref: signal/tracking.py:192-297.

    Both inputs z-scored (template over the ROI only, :251-252); template embedded into a
    float32 zero frame at slices_yx (:254-260); cross-power spectrum whitened with
    prod / (|prod| + eps) (:280-281); |fftshift(ifft2(.))| searched with first-occurrence
    argmax (:283-286); peak = map value at the maximum, snr = |peak| / (median(|map|) + eps)
    (:314-321); shifts relative to (H//2, W//2) plus the (swapped) Taylor terms (:289-295).
    """
    tpl = _float2d(template, "template")
    img = _float2d(image, "image")
    H, W = img.shape
    if slices_yx is None:
        slices_yx = centered_roi_slices((H, W), tpl.shape)
    sy, sx = slices_yx
    if tpl.shape != (sy.stop - sy.start, sx.stop - sx.start):
        raise ValueError("ROI shape does not match target slice dimensions.")
    img_z = zscore2d(img, eps)
    pad = np.zeros((H, W), dtype=np.float32)
    pad[sy, sx] = zscore2d(tpl, eps)
    prod = np.fft.fft2(img_z) * np.conj(np.fft.fft2(pad))
    mag = np.abs(np.fft.fftshift(np.fft.ifft2(prod / (np.abs(prod) + eps))))
    i, j = np.unravel_index(np.argmax(mag), mag.shape)
    peak = float(mag[i, j])
    snr = float(abs(peak) / (float(np.median(np.abs(mag))) + eps))
    dy, dx = float(i - H // 2), float(j - W // 2)
    if subpixel:
        di, dj = peak_subpixel_taylor(mag, int(i), int(j))
        dy += di
        dx += dj
    return dy, dx, peak, snr


# --------------------------------------------------------------------------------------
# metrics.statistics / metrics.sharpness
# --------------------------------------------------------------------------------------

def distribution_moments(image, saturation_value=65535.0, eps=1e-6):
    """Moments over finite pixels in float64.  ref: metrics/statistics.py:17-125.

    skewness = m3/m2^1.5 and kurtosis = m4/m2^2 - 3 (scipy.stats.describe, biased, :79-81);
    frac_zero = mean(|x| <= eps) (:83); SNRdB = 20 log10(mean/std) with its edge cases (:85-94);
    frac_sat = mean(x >= saturation_value) or NaN (:96-99).
    """
    data = np.asarray(image)
    if data.ndim not in (1, 2):
        raise ValueError(f"Expected 1D or 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError("distribution_moments received an empty image.")
    x = np.asarray(data, dtype=np.float64).ravel()
    x = x[np.isfinite(x)]
    if x.size == 0:
        raise ValueError("distribution_moments received no finite values.")
    mean = float(x.mean())
    std = float(x.std(ddof=0))
    d = describe(x, axis=None)
    if std == 0.0:
        snr_db = float("inf") if mean > 0.0 else float("nan")
    else:
        q = mean / std
        snr_db = float(20.0 * np.log10(q)) if q > 0.0 else (float("-inf") if q == 0.0 else float("nan"))
    return {
        "mean": mean,
        "std": std,
        "variance": float(std * std),
        "skewness": float(d.skewness),
        "kurtosis": float(d.kurtosis),
        "frac_zero": float(np.mean(np.abs(x) <= eps)),
        "frac_sat": float("nan") if saturation_value is None else float(np.mean(x >= float(saturation_value))),
        "SNRdB": snr_db,
    }


def tenengrad(image, eps=1e-12):
    """Mean squared Sobel gradients (reflect borders) over finite pixels.  ref: metrics/sharpness.py:405-476."""
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError("tenengrad received an empty image.")
    finite = np.isfinite(data)
    if not finite.any():
        raise ValueError("tenengrad received image with no finite values.")
    x = np.asarray(data, dtype=float)
    gx = ndimage.sobel(x, axis=1, mode="reflect")
    gy = ndimage.sobel(x, axis=0, mode="reflect")
    ex = float(np.mean((gx * gx)[finite]))
    ey = float(np.mean((gy * gy)[finite]))
    return {"tenengrad": float(ex + ey), "ex": ex, "ey": ey, "re": float(ex / (ey + float(eps)))}


def laplacian_variance(image):
    """Population variance of the 5-point Laplacian (reflect borders).  ref: metrics/sharpness.py:482-530."""
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError("laplacian_variance received an empty image.")
    finite = np.isfinite(data)
    if not finite.any():
        raise ValueError("laplacian_variance received image with no finite values.")
    lap = ndimage.laplace(np.asarray(data, dtype=float), mode="reflect")
    return float(np.var(lap[finite], ddof=0))


def sobel_reflect(x):
    """Explicit restatement of scipy.ndimage.sobel(mode="reflect") for both axes.

    gx[i,j] = sum_{d in -1,0,1} w_d (x[i+d, j+1] - x[i+d, j-1]),  w = (1, 2, 1); gy likewise
    with the roles of the axes exchanged; "reflect" duplicates the edge sample (numpy "symmetric").
    """
    p = np.pad(np.asarray(x, dtype=float), 1, mode="symmetric")
    c = slice(1, -1)
    gx = ((p[:-2, 2:] - p[:-2, :-2]) + 2.0 * (p[c, 2:] - p[c, :-2]) + (p[2:, 2:] - p[2:, :-2]))
    gy = ((p[2:, :-2] - p[:-2, :-2]) + 2.0 * (p[2:, c] - p[:-2, c]) + (p[2:, 2:] - p[:-2, 2:]))
    return gx, gy


def laplace_reflect(x):
    """Explicit restatement of scipy.ndimage.laplace(mode="reflect")."""
    p = np.pad(np.asarray(x, dtype=float), 1, mode="symmetric")
    return p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4.0 * p[1:-1, 1:-1]


# --------------------------------------------------------------------------------------
# geometry / maths helpers used by grain, bandwidth, inverse_autocorr_width
# --------------------------------------------------------------------------------------

def pad_to_square(image, fill_value=0.0):
    """Centre a (H, W) array in an (N, N) one, N = max(H, W), y0 = (N-H)//2.  ref: geometry/masks.py:11-56."""
    a = np.asarray(image)
    H, W = a.shape
    N = max(H, W)
    out = np.full((N, N), fill_value, dtype=a.dtype)
    y0, x0 = (N - H) // 2, (N - W) // 2
    out[y0:y0 + H, x0:x0 + W] = a
    return out


def width_at_fraction(profile, fraction=INV_E, center_index=None):
    """Full width of a peak at fraction*peak, linear interpolation at both crossings.  ref: maths/stats.py:9-89.

    Strict '<' tests; returns (len(profile), True) when a side never crosses (quirk 11).
    """
    p = np.asarray(profile, dtype=float)
    if p.ndim != 1 or p.size == 0:
        raise ValueError("profile must be a non-empty 1D array.")
    if not (0.0 < fraction < 1.0):
        raise ValueError("fraction must be in (0, 1).")
    c = int(np.argmax(p)) if center_index is None else int(center_index)
    c = max(0, min(c, p.size - 1))
    thr = p[c] * fraction
    below_l = np.nonzero(p[:c + 1] < thr)[0]
    below_r = np.nonzero(p[c:] < thr)[0]
    if below_l.size == 0 or below_r.size == 0:
        return float(p.size), True
    il = int(below_l[-1])
    ir = c + int(below_r[0])
    ya, yb = p[il], p[il + 1]
    xl = float(il) if yb == ya else il + (thr - ya) / (yb - ya)
    ya, yb = p[ir - 1], p[ir]
    xr = float(ir) if yb == ya else (ir - 1) + (thr - ya) / (yb - ya)
    return float(xr - xl), False


def distance_at_fraction_from_peak(profile, fraction=INV_E, peak_index=0):
    """One-sided distance to the first sample below fraction*peak.  ref: maths/stats.py:92-155."""
    p = np.asarray(profile, dtype=float)
    if p.ndim != 1 or p.size == 0:
        raise ValueError("profile must be a non-empty 1D array.")
    if not (0.0 < fraction < 1.0):
        raise ValueError("fraction must be in (0, 1).")
    k0 = max(0, min(int(peak_index), p.size - 1))
    thr = p[k0] * fraction
    below = np.nonzero(p[k0:] < thr)[0]
    if below.size == 0:
        return float(p.size), True
    ic = k0 + int(below[0])
    if ic == k0:
        return 0.0, False
    ya, yb = p[ic - 1], p[ic]
    xc = float(ic) if yb == ya else (ic - 1) + (thr - ya) / (yb - ya)
    return float(xc - k0), False


def radial_mean_interpolated(z, fill_value=0.0):
    """Angular mean of a bilinear polar resampling about (ny//2, nx//2).  ref: maths/radial.py:101-169.

    r_max = min(max|x|, max|y|) with x = arange(n) - n//2 (:142-143); nr = floor(r_max)+1 samples
    on linspace(0, r_max) (:147-148, :157); ntheta = int(2*pi*180) = 1130 (:150); sample at
    (y, x) = (r sin(theta), r cos(theta)) (:160-166), 0 outside the grid.
    """
    z = np.asarray(z, dtype=float)
    if z.ndim != 2:
        raise ValueError("signal_2d must be a 2D array.")
    if not np.isfinite(z).all():
        raise ValueError("signal_2d contains non-finite values.")
    ny, nx = z.shape
    x = np.arange(nx, dtype=float) - (nx // 2)
    y = np.arange(ny, dtype=float) - (ny // 2)
    r_max = min(float(np.max(np.abs(x))), float(np.max(np.abs(y))))
    if r_max <= 0:
        raise ValueError("r_max must be > 0")
    nr = int(np.floor(r_max)) + 1
    ntheta = int(2.0 * np.pi * 180.0)
    if nr <= 1:
        raise ValueError("nr must be > 1.")
    r = np.linspace(0.0, r_max, nr)
    theta = np.linspace(0.0, 2.0 * np.pi, ntheta, endpoint=False)
    R, TH = np.meshgrid(r, theta, indexing="ij")
    interp = RegularGridInterpolator((y, x), z, bounds_error=False, fill_value=fill_value)
    vals = interp(np.column_stack([(R * np.sin(TH)).ravel(), (R * np.cos(TH)).ravel()])).reshape(R.shape)
    return vals.mean(axis=1), r


def bilinear_polar_mean(z):
    """Explicit restatement of radial_mean_interpolated's sampling (what the CUDA kernel evaluates).

    For each (r_k, theta_m): (yy, xx) = (cy + r sin, cx + r cos) in index units; samples outside
    [0, n-1] contribute 0; inside, the cell index is clamped to n-2 and the four corners are
    blended with weights (1-ty)(1-tx) etc.  (scipy's linear RegularGridInterpolator on a unit grid).
    """
    z = np.asarray(z, dtype=float)
    ny, nx = z.shape
    cy, cx = ny // 2, nx // 2
    r_max = float(min(max(cx, nx - 1 - cx), max(cy, ny - 1 - cy)))
    nr = int(np.floor(r_max)) + 1
    ntheta = int(2.0 * np.pi * 180.0)
    r = np.linspace(0.0, r_max, nr)
    theta = np.linspace(0.0, 2.0 * np.pi, ntheta, endpoint=False)
    R, TH = np.meshgrid(r, theta, indexing="ij")
    yy = R * np.sin(TH) + cy
    xx = R * np.cos(TH) + cx
    inside = (yy >= 0) & (yy <= ny - 1) & (xx >= 0) & (xx <= nx - 1)
    iy = np.clip(np.floor(yy).astype(np.int64), 0, ny - 2)
    ix = np.clip(np.floor(xx).astype(np.int64), 0, nx - 2)
    ty, tx = yy - iy, xx - ix
    v = (z[iy, ix] * (1 - ty) * (1 - tx) + z[iy, ix + 1] * (1 - ty) * tx
         + z[iy + 1, ix] * ty * (1 - tx) + z[iy + 1, ix + 1] * ty * tx)
    return np.where(inside, v, 0.0).mean(axis=1), r


def percentile_linear(x, q):
    """numpy's default 'linear' percentile on finite data: h = q/100 (n-1), s[fl] + (h-fl)(s[fl+1]-s[fl])."""
    s = np.sort(np.asarray(x, dtype=float).ravel())
    h = q / 100.0 * (s.size - 1)
    lo = int(math.floor(h))
    hi = min(lo + 1, s.size - 1)
    return float(s[lo] + (h - lo) * (s[hi] - s[lo]))


# --------------------------------------------------------------------------------------
# metrics.speckles
# --------------------------------------------------------------------------------------

def amplitude(image):
    """visibility = nanstd/nanmean, contrast from the 0.05 / 99.95 percentiles.  ref: metrics/speckles.py:602-663, utils/range.py:44-54."""
    img = np.asarray(image, dtype=float)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    mu = float(np.nanmean(img))
    if not np.isfinite(mu) or mu <= 0.0:
        raise ValueError("Mean intensity must be positive and finite.")
    vis = float(np.nanstd(img)) / mu
    vmin = float(np.nanpercentile(img, 0.05))
    vmax = float(np.nanpercentile(img, 99.95))
    den = vmax + vmin
    if not np.isfinite(den) or den <= 0.0:
        raise ValueError("Invalid percentile range for Michelson contrast.")
    return {"visibility": vis, "contrast": (vmax - vmin) / den}


def _autocorr_widths(ac, fraction):
    """Shared tail of grain / inverse_autocorr_width: cuts through argmax + radial 1/e distance."""
    iy, ix = np.unravel_index(int(np.argmax(ac)), ac.shape)
    ly, _ = width_at_fraction(ac[:, ix], fraction=fraction, center_index=iy)
    lx, _ = width_at_fraction(ac[iy, :], fraction=fraction, center_index=ix)
    rad, r = radial_mean_interpolated(ac)
    dr = float(r[1] - r[0])
    dist, _ = distance_at_fraction_from_peak(rad, fraction=fraction, peak_index=0)
    return float(lx), float(ly), float(2.0 * float(dist) * dr)


def grain(image, fraction=INV_E):
    """1/e widths of the peak-normalised autocorrelation.  ref: metrics/speckles.py:497-596."""
    data = np.asarray(image, dtype=float)
    if data.ndim != 2:
        raise ValueError("image must be a 2D array.")
    if min(data.shape) < 128:
        raise ValueError("image too small for speckle grain metrics (min dimension < 128).")
    data = pad_to_square(data, fill_value=np.mean(data))
    ac, xl, yl = autocorr2d(data, remove_mean=True, standardize=False, normalize="peak")
    lx, ly, leq = _autocorr_widths(ac, fraction)
    return {"lx": lx, "ly": ly, "leq": leq, "r": float(lx / ly) if ly != 0 else float("inf"),
            "autocorr": np.asarray(ac, dtype=float), "xlag": xl, "ylag": yl}


def inverse_autocorr_width(image, fraction=INV_E, min_size_px=32):
    """Inverse 1/e widths of the standardised autocorrelation.  ref: metrics/sharpness.py:635-746."""
    data = np.asarray(image, dtype=float)
    if data.ndim != 2:
        raise ValueError("image must be a 2D array.")
    if data.size == 0:
        raise ValueError("inverse_autocorr_width received an empty image.")
    if min(data.shape) < int(min_size_px):
        raise ValueError("image too small for inverse autocorrelation width")
    data = pad_to_square(data, fill_value=np.mean(data))
    ac, _, _ = autocorr2d(data, remove_mean=True, standardize=True, normalize="peak")
    lx, ly, leq = _autocorr_widths(ac, fraction)
    inv = lambda v: float(1.0 / v) if v != 0.0 else float("inf")
    return {"sx": inv(lx), "sy": inv(ly), "seq": inv(leq), "r": float(lx / ly) if ly != 0.0 else float("inf")}


def bandwidth(image):
    """PSD second moments, f95 and participation ratio inside the inscribed circle.  ref: metrics/speckles.py:669-817."""
    img = np.asarray(image, dtype=float)
    if img.ndim != 2:
        raise ValueError("image must be a 2D array.")
    img = pad_to_square(img, fill_value=np.mean(img))
    mu = float(np.nanmean(img))
    if not np.isfinite(mu):
        raise ValueError("image mean is not finite.")
    P, fx, fy = psd2d(img - mu, scale=True)
    P = np.nan_to_num(np.asarray(P, dtype=float), nan=0.0, posinf=0.0, neginf=0.0).copy()
    ny, nx = P.shape
    P[ny // 2, nx // 2] = 0.0
    FX, FY = np.meshgrid(fx, fy, indexing="xy")
    FR = np.sqrt(FX * FX + FY * FY)
    keep = FR <= min(float(np.max(np.abs(fx))), float(np.max(np.abs(fy))))
    Pm, FXm, FYm, FRm = P[keep], FX[keep], FY[keep], FR[keep]
    total = float(Pm.sum())
    if not np.isfinite(total) or total <= 0.0:
        raise ValueError("PSD energy is not positive/finite after mean/DC removal.")
    feq = float(np.sqrt(np.sum(FRm * FRm * Pm) / total))
    sfx = float(np.sqrt(np.sum(FXm * FXm * Pm) / total))
    sfy = float(np.sqrt(np.sum(FYm * FYm * Pm) / total))
    order = np.argsort(FRm)
    cdf = np.cumsum(Pm[order]) / total
    idx = min(int(np.searchsorted(cdf, 0.95, side="left")), FRm.size - 1)
    p = Pm / total
    return {"feq": feq, "f95": float(FRm[order][idx]), "sig_fx": sfx, "sig_fy": sfy,
            "rf": float(sfx / sfy) if sfy != 0.0 else float("inf"), "spr": float(1.0 / float(np.sum(p * p)))}


def eigenvalues(image, k=5, eps=1e-30):
    """STA2: sum of the k leading eigenvalues of S = J J^T / (M N - 1), J the energy-normalised, mean-removed image;
    eigenvalues from the singular values of J.  ref: metrics/sharpness.py:752-861 (checks :811-837, svd :839-847)."""
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError("eigenvalues received an empty image.")
    if not np.all(np.isfinite(data)):
        raise ValueError("eigenvalues requires all values to be finite.")
    if int(k) < 1:
        raise ValueError("k must be >= 1.")
    x = np.asarray(data, dtype=float)
    energy = float(np.sqrt(np.sum(x * x)))
    if not np.isfinite(energy) or energy <= 0.0:
        raise ValueError("eigenvalues cannot normalize an all-zero image.")
    J = x / energy
    J = J - float(np.mean(J))
    if J.size < 2:
        raise ValueError("eigenvalues requires at least 2 pixels (M*N >= 2).")
    sv = np.linalg.svd(J, full_matrices=False, compute_uv=False)
    eig = sv * sv / float(J.size - 1)
    e1, e2 = float(eig[0]), float(eig[1]) if eig.size >= 2 else 0.0
    return {"eigenvalues": float(np.sum(eig[:min(int(k), eig.size)])), "e1": e1, "e2": e2, "re": float(e1 / (e2 + float(eps)))}


def spectral_entropy(image, eps=1e-30):
    """Normalised Shannon entropy of the unscaled PSD, DC excluded.  ref: metrics/sharpness.py:536-629.

    The pad_to_square result at :590 is discarded at :591, so no padding happens (quirk 9).
    """
    data = np.asarray(image)
    if data.ndim != 2:
        raise ValueError(f"Expected 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError("spectral_entropy received an empty image.")
    if not np.all(np.isfinite(data)):
        raise ValueError("spectral_entropy requires all values to be finite.")
    x = np.asarray(data, dtype=float)
    P, _, _ = psd2d(x - float(x.mean()), scale=False)
    P = np.asarray(P, dtype=float).copy()
    P[P.shape[0] // 2, P.shape[1] // 2] = 0.0
    s = float(P.sum())
    if not np.isfinite(s) or s <= 0.0:
        raise ValueError("PSD sum is non-positive; cannot compute spectral entropy.")
    p = np.clip(P.ravel() / s, float(eps), None)
    return float(-np.sum(p * np.log(p)) / np.log(float(p.size - 1)))


# --------------------------------------------------------------------------------------
# preprocessing.normalize
# --------------------------------------------------------------------------------------

def flat_field_correction(images, flats=None, darks=None, scale="flat_median", eps=None, bad_pixel_removal=False):
    """(I - D) / (F - D) * s in float32 with a bad-pixel mask (no median repair).  ref: preprocessing/normalize.py:12-145."""
    if scale not in {"none", "flat_mean", "flat_median"}:
        raise ValueError(f"Invalid scale option: {scale}")
    img = np.asarray(images).astype(np.float32, copy=False)
    if img.ndim not in (2, 3):
        raise ValueError("images must be 2D or 3D")

    def collapse(a):
        if a is None:
            return None
        a = np.asarray(a)
        if a.ndim == 3:
            return a.astype(np.float32).mean(axis=0)
        if a.ndim == 2:
            return a.astype(np.float32)
        raise ValueError("flats/darks must be 2D or 3D")

    F, D = collapse(flats), collapse(darks)
    if F is None and D is None:
        return img.copy()
    if D is None:
        D = np.zeros_like(F)
    if F is None:
        return img - D
    den = F - D
    if eps is None:
        med = np.median(den)
        eps = 1e-6 * med if med > 0 else 1e-6
    bad = den <= eps
    safe = den.copy()
    safe[bad] = 1.0
    out = (img - D) / safe
    if scale != "none":
        out *= np.mean(den[~bad]) if scale == "flat_mean" else np.median(den[~bad])
    out[..., bad] = 0.0
    if bad_pixel_removal:
        # ref: preprocessing/normalize.py:134-140 -- 3x3 median (scipy default 'reflect') of the frame with zeroed bad pixels
        rep = ndimage.median_filter(out, size=(1, 3, 3) if out.ndim == 3 else (3, 3))
        out[..., bad] = rep[..., bad]
    return out.astype(np.float32, copy=False)


# --------------------------------------------------------------------------------------
# row T: per-pixel temporal moments (no reference function -> parity unpinned)
# --------------------------------------------------------------------------------------

def temporal_moments(stack):
    """Per-pixel mean / std(ddof 0) / variance / skewness / excess kurtosis over axis 0, float64.

    PARITY UNPINNED: the reference has no per-pixel temporal-moment function (SURVEY.md 8(a)
    row T); this lifts distribution_moments' definitions (statistics.py:75-81) along the
    time axis with the same scipy.stats.describe call.
    """
    x = np.asarray(stack, dtype=np.float64)
    if x.ndim != 3:
        raise ValueError("stack must be (T, H, W)")
    d = describe(x, axis=0)
    std = x.std(axis=0)
    return {"mean": x.mean(axis=0), "std": std, "variance": std * std,
            "skewness": np.asarray(d.skewness), "kurtosis": np.asarray(d.kurtosis)}


# --------------------------------------------------------------------------------------
# tiling executor (metrics/common.py)
# --------------------------------------------------------------------------------------

def split_edges(length, n_parts):
    """ref: metrics/common.py:75-106 -- edges at round(linspace(0, length, n_parts + 1)), last stop = length."""
    edges = np.linspace(0, length, n_parts + 1)
    out = []
    for i in range(n_parts):
        a = int(round(float(edges[i])))
        b = max(int(round(float(edges[i + 1]))), a + 1)
        out.append((a, b))
    out[-1] = (out[-1][0], length)
    return out


def choose_tiling_mode(h, w, min_tile_px=128):
    """ref: metrics/common.py:109-170 (tiles=True branch)."""
    if h // 9 >= min_tile_px and w // 9 >= min_tile_px:
        return "subtiles_9x9"
    if h // 3 >= min_tile_px and w // 3 >= min_tile_px:
        return "tiles_3x3"
    return "off"


def tiled_scalar_fields(image, tile_mode, compute_fn):
    """ref: metrics/common.py:278-378 -- {field: {"mean": 3x3, "std": 3x3}}; tiles_3x3: values and NaN; subtiles_9x9:
    np.mean / np.std(ddof=0) of each 3x3 block of the 9x9 sub-tile grid (:248-275)."""
    img = np.asarray(image)
    h, w = img.shape
    n = 3 if tile_mode == "tiles_3x3" else 9
    ye, xe = split_edges(h, n), split_edges(w, n)
    grids = {}
    for r, (y0, y1) in enumerate(ye):
        for c, (x0, x1) in enumerate(xe):
            for k, v in compute_fn(img[y0:y1, x0:x1]).items():
                grids.setdefault(k, np.empty((n, n)))[r, c] = float(v)
    out = {}
    for k, g in grids.items():
        if n == 3:
            out[k] = {"mean": g, "std": np.full((3, 3), np.nan)}
        else:
            mean, std = np.empty((3, 3)), np.empty((3, 3))
            for r in range(3):
                for c in range(3):
                    blk = g[3 * r:3 * r + 3, 3 * c:3 * c + 3]
                    mean[r, c], std[r, c] = np.mean(blk), np.std(blk, ddof=0)
            out[k] = {"mean": mean, "std": std}
    return out


def speckle_tiles(image, display_origin="lower", saturation_value=65535.0, eps=1e-6):
    """ref: metrics/speckles.py:148,192-250 -- the "tiles" block of speckle_stats(image, tiles=True), all groups."""
    img = np.asarray(image)[::-1, :] if display_origin == "lower" else np.asarray(image)
    mode = choose_tiling_mode(*img.shape)
    scal = lambda d, keys: {k: float(d[k]) for k in keys}
    return mode, {
        "amplitude": tiled_scalar_fields(img, mode, lambda t: scal(amplitude(t), ("visibility", "contrast"))),
        "grain": tiled_scalar_fields(img, mode, lambda t: scal(grain(t), ("lx", "ly", "leq", "r"))),
        "stats": tiled_scalar_fields(img, mode, lambda t: distribution_moments(t, saturation_value=saturation_value, eps=eps)),
        "bandwidth": tiled_scalar_fields(img, mode, lambda t: scal(bandwidth(t), ("spr", "feq", "f95", "sig_fx", "sig_fy", "rf"))),
    }


def sharpness_tiles(image, display_origin="lower", saturation_value=65535.0, eps=1e-6):
    """ref: metrics/sharpness.py:163,213-282 -- the "tiles" block of sharpness_stats(image, tiles=True) without eigenvalues."""
    img = np.asarray(image)[::-1, :] if display_origin == "lower" else np.asarray(image)
    mode = choose_tiling_mode(*img.shape)
    scal = lambda d, keys: {k: float(d[k]) for k in keys}
    return mode, {
        "stats": tiled_scalar_fields(img, mode, lambda t: distribution_moments(t, saturation_value=saturation_value, eps=eps)),
        "gradient": tiled_scalar_fields(img, mode, lambda t: scal(tenengrad(t), ("tenengrad", "ex", "ey", "re"))),
        "laplacian": tiled_scalar_fields(img, mode, lambda t: {"laplacian_variance": laplacian_variance(t)}),
        "spectral": tiled_scalar_fields(img, mode, lambda t: {"spectral_entropy": spectral_entropy(t)}),
        "autocorrelation": tiled_scalar_fields(img, mode, lambda t: scal(inverse_autocorr_width(t), ("sx", "sy", "seq", "r"))),
    }


# --------------------------------------------------------------------------------------
# signal.tracking.template_matching (opencv backend: cv2.matchTemplate TM_CCOEFF_NORMED)
# --------------------------------------------------------------------------------------

def ncc_valid(img, tpl):
    """cv2.matchTemplate(img, tpl, TM_CCOEFF_NORMED), restated: result (H-h+1, W-w+1) float32.

    OpenCV (modules/imgproc/src/templmatch.cpp, common_matchTemplate; opencv-python is an unpinned dependency of the
    reference, 4.13.0 where the goldens were recorded): numerator = cross-correlation - window_sum * mean(tpl);
    denominator = sqrt(max(window_sqsum - window_sum^2 / area, 0)) * std(tpl) * sqrt(area), window sums from
    double-precision integral images; |num| < den -> num / den, |num| < 1.125 den -> +-1, else 0."""
    I = np.asarray(img, dtype=np.float64)
    T = np.asarray(tpl, dtype=np.float64)
    H, W = I.shape
    h, w = T.shape
    area = float(h * w)
    Tz = T - T.mean()
    fi = np.fft.rfft2(I)
    ft = np.fft.rfft2(Tz, s=(H, W))
    num = np.fft.irfft2(fi * np.conj(ft), s=(H, W))[:H - h + 1, :W - w + 1]
    def box(a):
        P = np.zeros((H + 1, W + 1))
        P[1:, 1:] = a.cumsum(0).cumsum(1)
        return P[h:, w:] - P[:-h, w:] - P[h:, :-w] + P[:-h, :-w]
    s1, s2 = box(I), box(I * I)
    diff2 = np.maximum(s2 - s1 * s1 / area, 0.0)
    t = np.where(diff2 <= np.minimum(0.5, 10 * np.finfo(np.float32).eps * s2), 0.0, np.sqrt(diff2) * T.std() * math.sqrt(area))
    out = np.where(np.abs(num) < t, num / np.where(t > 0, t, 1.0), np.where(np.abs(num) < 1.125 * t, np.sign(num), 0.0))
    return out.astype(np.float32)


def template_matching(template, image, slices_yx=None, subpixel=True, eps=1e-9):
    """ref: signal/tracking.py:82-188 (backend="opencv")."""
    tpl, img = _float2d(template, "template"), _float2d(image, "image")
    H, W = img.shape
    h, w = tpl.shape
    if h > H or w > W:
        raise ValueError(f"template shape {(h, w)} must fit inside image shape {(H, W)}")
    if slices_yx is None:
        slices_yx = centered_roi_slices((H, W), (h, w))
    sy, sx = slices_yx
    y0, x0 = (sy.start + sy.stop - 1) / 2.0, (sx.start + sx.stop - 1) / 2.0
    tz = zscore2d(tpl, eps).astype(np.float32)
    iz = zscore2d(img, eps).astype(np.float32)
    corr = ncc_valid(iz, tz)
    i, j = np.unravel_index(int(np.argmax(corr)), corr.shape)
    peak = float(corr[i, j])
    snr = float(abs(peak) / (float(np.median(np.abs(corr))) + eps))
    py, px = float(i), float(j)
    if subpixel:
        di, dj = peak_subpixel_taylor(corr, i, j)
        py += float(di)
        px += float(dj)
    return float(py + (h - 1) / 2.0 - y0), float(px + (w - 1) / 2.0 - x0), peak, snr
