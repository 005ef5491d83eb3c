"""
Stack-level entry points: thin, typed wrappers over the C ABI that work on HBM-resident stacks.

Every function takes a (T, ny, nx) float32 CUDA tensor (``as_stack`` uploads numpy input once)
and returns either small numpy tables (one row per frame) or CUDA tensors for maps, so that
stacks and maps stay in HBM between stages.  The per-frame drop-in functions in
``barc4dip_b200.signal`` / ``.metrics`` / ``.preprocessing`` are T = 1 calls of these.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import FR, FR_NCOLS, SP, SP_NCOLS, get_context, ptr, require_cuda

__all__ = [
    "as_stack", "frame_reductions", "select_quantiles", "flat_field", "flat_gain",
    "temporal_moments", "TemporalAccumulator", "fft2d", "psd2d", "autocorr2d", "xcorr2d",
    "PhaseTracker", "stack_pipeline", "check_fft_shape", "bad_pixel_repair", "template_match",
]


def as_stack(a, device: int | None = None):
    """(ny, nx) or (T, ny, nx) array-like -> contiguous float32 CUDA tensor of shape (T, ny, nx)."""
    t = _lib.as_device_f32(a, device)
    if t.ndim == 2:
        t = t.unsqueeze(0)
    if t.ndim != 3:
        raise ValueError(f"expected a 2D frame or a 3D stack, got ndim={t.ndim}")
    return t


def _dev(t) -> int:
    return t.device.index if t.device.index is not None else 0


DFT_MAX = 4096      # B4D_DFT_MAX: largest side of the chirp-z path


def check_fft_shape(ny: int, nx: int, generic_ok: bool = False):
    """Power-of-two sides in [128, 2048] run the hot-path kernels. Entry points with generic_ok also take any sides in
    [1, 4096] through the Bluestein path (csrc/generic_dft.cuh): tiles, odd detector formats, frames wider than 2048."""
    pow2 = all(128 <= n <= _lib.FFT_MAX and (n & (n - 1)) == 0 for n in (ny, nx))
    if pow2 or (generic_ok and all(1 <= n <= DFT_MAX for n in (ny, nx))):
        return
    raise _lib.B4DUnsupported(
        f"the sm_100a FFT kernels cover power-of-two sides in [128, {_lib.FFT_MAX}]"
        + (f" or any sides in [1, {DFT_MAX}]" if generic_ok else "")
        + f"; got (ny, nx) = ({ny}, {nx}). There is no CPU fallback on this path.")


# ------------------------------------------------------------------------------------------
# reductions / selection
# ------------------------------------------------------------------------------------------

def frame_reductions(stack, *, gain=None, dark=None, saturation_value: float | None = 65535.0,
                     eps: float = 1e-6, return_device: bool = False):
    """Per-frame single-pass reductions -> float64 table (T, FR_NCOLS) (columns: ``_lib.FR``)."""
    torch = require_cuda()
    T, ny, nx = stack.shape
    ctx = get_context(_dev(stack))
    out = torch.empty((T, FR_NCOLS), dtype=torch.float64, device=stack.device)
    sat = float("nan") if saturation_value is None else float(saturation_value)
    ctx.check(ctx.lib.b4d_frame_reductions(ctx.handle, ptr(stack), T, ny, nx, ptr(gain), ptr(dark),
                                           sat, float(eps), ptr(out)), "b4d_frame_reductions")
    return out if return_device else out.cpu().numpy()


def select_quantiles(stack, quantiles, *, use_abs: bool = False, return_device: bool = False):
    """Exact order statistics bracketing each quantile.

    Returns (values (T, 2*len(q)) float32, n_valid (T,) int64): for quantile q the two columns hold
    sorted[floor(h)] and sorted[min(floor(h)+1, n-1)] with h = numpy's 'linear' virtual index.
    """
    torch = require_cuda()
    T = stack.shape[0]
    n = int(np.prod(stack.shape[1:]))
    q = np.ascontiguousarray(np.asarray(quantiles, dtype=np.float64).ravel())
    ctx = get_context(_dev(stack))
    out = torch.empty((T, 2 * q.size), dtype=torch.float32, device=stack.device)
    nv = torch.empty((T,), dtype=torch.int64, device=stack.device)
    ctx.check(ctx.lib.b4d_select_ranks(ctx.handle, ptr(stack), T, n, q.ctypes.data_as(C.c_void_p), int(q.size),
                                       int(bool(use_abs)), ptr(out), ptr(nv)), "b4d_select_ranks")
    if return_device:
        return out, nv
    return out.cpu().numpy(), nv.cpu().numpy()


def virtual_index(n: int, q: float) -> float:
    """numpy's 'linear' quantile index, same expression as numpy (and as the device code)."""
    return n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0


def lerp_like_numpy(a: float, b: float, t: float) -> float:
    """numpy.lib._function_base_impl._lerp: a + (b-a)t, switched to b - (b-a)(1-t) for t >= 0.5."""
    d = b - a
    return b - d * (1.0 - t) if t >= 0.5 else a + d * t


def quantile_from_bracket(lo: float, hi: float, n: int, q: float) -> float:
    h = virtual_index(n, q)
    g = h - math.floor(h)
    return lerp_like_numpy(float(lo), float(hi), g)


# ------------------------------------------------------------------------------------------
# flat field / temporal moments
# ------------------------------------------------------------------------------------------

def flat_gain(flat, dark, *, eps: float, scale_value: float):
    torch = require_cuda()
    ny, nx = flat.shape[-2:]
    ctx = get_context(_dev(flat))
    gain = torch.empty((ny, nx), dtype=torch.float32, device=flat.device)
    ctx.check(ctx.lib.b4d_flat_gain(ctx.handle, ptr(flat), ptr(dark), ny, nx, float(eps), float(scale_value),
                                    ptr(gain)), "b4d_flat_gain")
    return gain


def flat_field(stack, flat, dark, *, eps: float, scale_value: float, apply_scale: bool):
    torch = require_cuda()
    T, ny, nx = stack.shape
    ctx = get_context(_dev(stack))
    out = torch.empty_like(stack)
    ctx.check(ctx.lib.b4d_flat_field(ctx.handle, ptr(stack), T, ny, nx, ptr(flat), ptr(dark), float(eps),
                                     float(scale_value), int(bool(apply_scale)), ptr(out)), "b4d_flat_field")
    return out


def bad_pixel_repair(frames, flat, dark, *, eps: float):
    """In place: frames[:, bad] = 3x3 median (reflect border) of the flat-field-corrected frames (b4d_bad_pixel_repair)."""
    T, ny, nx = frames.shape
    ctx = get_context(_dev(frames))
    ctx.check(ctx.lib.b4d_bad_pixel_repair(ctx.handle, ptr(frames), T, ny, nx, ptr(flat), ptr(dark), float(eps)),
              "b4d_bad_pixel_repair")
    return frames


def sub(a, b):
    torch = require_cuda()
    ctx = get_context(_dev(a))
    out = torch.empty_like(a)
    ctx.check(ctx.lib.b4d_sub(ctx.handle, ptr(a), ptr(b), a.numel(), ptr(out)), "b4d_sub")
    return out


class TemporalAccumulator:
    """Streaming per-pixel moments over time; chunks may arrive one at a time (HBM-sized stacks).

    ``shift`` must be identical on every rank when the sums are all-reduced (see parallel.py).
    """

    def __init__(self, ny: int, nx: int, *, device: int | None = None, gain=None, dark=None, shift=None):
        torch = require_cuda()
        self.dev = _lib.default_device() if device is None else int(device)
        self.ny, self.nx = int(ny), int(nx)
        self.gain, self.dark = gain, dark
        self.sums = torch.zeros((4, ny, nx), dtype=torch.float64, device=f"cuda:{self.dev}")
        self.shift = shift
        self.count = 0

    def pilot(self, stack, n_frames: int = 16):
        """Set the shift map to the mean of the first n_frames (corrected) frames of ``stack``."""
        torch = require_cuda()
        ctx = get_context(self.dev)
        n = min(int(n_frames), stack.shape[0])
        self.shift = torch.empty((self.ny, self.nx), dtype=torch.float32, device=stack.device)
        ctx.check(ctx.lib.b4d_temporal_pilot(ctx.handle, ptr(stack), n, self.ny, self.nx, ptr(self.gain),
                                             ptr(self.dark), ptr(self.shift)), "b4d_temporal_pilot")
        return self.shift

    def update(self, stack):
        if self.shift is None:
            self.pilot(stack)
        T, ny, nx = stack.shape
        if (ny, nx) != (self.ny, self.nx):
            raise ValueError("frame shape mismatch")
        ctx = get_context(self.dev)
        ctx.check(ctx.lib.b4d_temporal_accumulate(ctx.handle, ptr(stack), T, ny, nx, ptr(self.gain), ptr(self.dark),
                                                  ptr(self.shift), ptr(self.sums)), "b4d_temporal_accumulate")
        self.count += int(T)

    def finalize(self, n_total: int | None = None, return_device: bool = False):
        torch = require_cuda()
        n = self.count if n_total is None else int(n_total)
        ctx = get_context(self.dev)
        maps = torch.empty((5, self.ny, self.nx), dtype=torch.float64, device=self.sums.device)
        ctx.check(ctx.lib.b4d_temporal_finalize(ctx.handle, ptr(self.sums), ptr(self.shift), n, self.ny, self.nx,
                                                ptr(maps)), "b4d_temporal_finalize")
        if return_device:
            return maps
        m = maps.cpu().numpy()
        return {"mean": m[0], "std": m[1], "variance": m[2], "skewness": m[3], "kurtosis": m[4]}


def temporal_moments(stack, *, gain=None, dark=None, return_device: bool = False):
    """Per-pixel mean/std/variance/skewness/kurtosis over axis 0 of an HBM-resident stack."""
    T, ny, nx = stack.shape
    acc = TemporalAccumulator(ny, nx, device=_dev(stack), gain=gain, dark=dark)
    acc.update(stack)
    return acc.finalize(return_device=return_device)


# ------------------------------------------------------------------------------------------
# FFT family
# ------------------------------------------------------------------------------------------

def fft2d(stack):
    """fftshift(fft2(frame)) for every frame -> complex64 CUDA tensor (T, ny, nx)."""
    torch = require_cuda()
    T, ny, nx = stack.shape
    check_fft_shape(ny, nx, generic_ok=True)
    ctx = get_context(_dev(stack))
    out = torch.empty((T, ny, nx, 2), dtype=torch.float32, device=stack.device)
    ctx.check(ctx.lib.b4d_fft2d(ctx.handle, ptr(stack), T, ny, nx, ptr(out)), "b4d_fft2d")
    return torch.view_as_complex(out)


def ifft2d(spec):
    """ifft2(ifftshift(F)) for every shifted complex64 spectrum of a (T, ny, nx) CUDA tensor -> complex64 (T, ny, nx)."""
    torch = require_cuda()
    T, ny, nx = spec.shape
    if not all(1 <= n <= DFT_MAX for n in (ny, nx)):
        raise _lib.B4DUnsupported(f"ifft2d covers sides in [2, {DFT_MAX}]; got ({ny}, {nx})")
    ctx = get_context(_dev(spec))
    src = torch.view_as_real(spec.to(torch.complex64).contiguous())
    out = torch.empty((T, ny, nx, 2), dtype=torch.float32, device=spec.device)
    ctx.check(ctx.lib.b4d_ifft2d(ctx.handle, ptr(src), T, ny, nx, ptr(out)), "b4d_ifft2d")
    return torch.view_as_complex(out)


def psd2d(stack, *, scale_factor: float = 1.0, sub_mean: bool = False, zero_dc: bool = False,
          want_map: bool = True, want_spectral: bool = False):
    """Shifted |FFT|^2 * scale_factor per frame. Returns (psd or None, spectral table or None)."""
    torch = require_cuda()
    T, ny, nx = stack.shape
    check_fft_shape(ny, nx, generic_ok=True)
    ctx = get_context(_dev(stack))
    out = torch.empty((T, ny, nx), dtype=torch.float32, device=stack.device) if want_map else None
    spec = torch.zeros((T, SP_NCOLS), dtype=torch.float64, device=stack.device) if want_spectral else None
    ctx.check(ctx.lib.b4d_psd2d(ctx.handle, ptr(stack), T, ny, nx, float(scale_factor), int(bool(sub_mean)),
                                int(bool(zero_dc)), ptr(out), ptr(spec)), "b4d_psd2d")
    return out, (spec.cpu().numpy() if spec is not None else None)


def _std_divisors(stack):
    """Per-frame population standard deviation on the device, 1 where it is 0 (signal/corr.py:229-235 divides only by a
    positive std) -> float32 (T, 1, 1)."""
    torch = require_cuda()
    tab = frame_reductions(stack, saturation_value=None, return_device=True)
    s = tab[:, FR["m2"]].sqrt()
    return torch.where(s > 0, s, torch.ones_like(s)).to(torch.float32)[:, None, None]


def autocorr2d(stack, *, remove_mean: bool = True, standardize: bool = False, normalize_peak: bool = True,
               want_map: bool = True, want_grain: bool = False, fraction: float = 1.0 / math.e):
    """Shifted circular autocorrelation per frame (+ optional grain widths table (T, 4))."""
    torch = require_cuda()
    T, ny, nx = stack.shape
    check_fft_shape(ny, nx, generic_ok=True)
    if standardize and not normalize_peak and not remove_mean and want_map and not want_grain:
        # the one combination the library leaves to the host: correlation is bilinear, so the /std of both factors is
        # applied to the finished map
        out, _ = autocorr2d(stack, remove_mean=False, standardize=False, normalize_peak=False)
        sd = _std_divisors(stack)
        return out.div_(sd * sd), None
    ctx = get_context(_dev(stack))
    out = torch.empty((T, ny, nx), dtype=torch.float32, device=stack.device) if want_map else None
    grain = torch.empty((T, 4), dtype=torch.float64, device=stack.device) if want_grain else None
    ctx.check(ctx.lib.b4d_autocorr2d(ctx.handle, ptr(stack), T, ny, nx, int(bool(remove_mean)),
                                     int(bool(standardize)), int(bool(normalize_peak)), ptr(out), float(fraction),
                                     ptr(grain)), "b4d_autocorr2d")
    return out, (grain.cpu().numpy() if grain is not None else None)


def xcorr2d(a, b, *, remove_mean: bool = True, standardize: bool = False, normalize_peak: bool = True):
    torch = require_cuda()
    T, ny, nx = a.shape
    if tuple(b.shape) != (T, ny, nx):
        raise ValueError("a and b must have the same shape.")
    check_fft_shape(ny, nx, generic_ok=True)
    if standardize and not normalize_peak:
        # left to the host by b4d_xcorr2d: correlation is bilinear, the two /std factors scale the finished map
        out = xcorr2d(a, b, remove_mean=remove_mean, standardize=False, normalize_peak=False)
        return out.div_(_std_divisors(a) * _std_divisors(b))
    ctx = get_context(_dev(a))
    out = torch.empty((T, ny, nx), dtype=torch.float32, device=a.device)
    ctx.check(ctx.lib.b4d_xcorr2d(ctx.handle, ptr(a), ptr(b), T, ny, nx, int(bool(remove_mean)),
                                  int(bool(standardize)), int(bool(normalize_peak)), ptr(out)), "b4d_xcorr2d")
    return out


class PhaseTracker:
    """Phase-correlation tracker bound to one reference template. The tracker OWNS the conjugate spectrum of its template
    (b4d_phase_reference_create): any number of trackers may be alive on a device, none re-targets another."""

    def __init__(self, template, frame_shape, *, y0: int, x0: int, eps: float = 1e-9, device: int | None = None):
        self.dev = _lib.default_device() if device is None else int(device)
        self.ny, self.nx = int(frame_shape[0]), int(frame_shape[1])
        self._ref = None
        check_fft_shape(self.ny, self.nx, generic_ok=True)     # sides that are not powers of two: map-based chirp-z path
        tpl = _lib.as_device_f32(template, self.dev)
        if tpl.ndim != 2:
            raise ValueError("template must be a 2D array.")
        h, w = tpl.shape
        if y0 < 0 or x0 < 0 or y0 + h > self.ny or x0 + w > self.nx:
            raise ValueError("ROI exceeds image bounds.")
        self.eps = float(eps)
        ctx = get_context(self.dev)
        ref = C.c_void_p()
        ctx.check(ctx.lib.b4d_phase_reference_create(ctx.handle, ptr(tpl), h, w, self.ny, self.nx, int(y0), int(x0),
                                                     self.eps, C.byref(ref)), "b4d_phase_reference_create")
        self._ref = ref
        _last_tracker[self.dev] = self

    @property
    def handle(self):
        if self._ref is None or not self._ref.value:
            raise _lib.B4DError("this PhaseTracker has been closed")
        return self._ref

    def close(self):
        """Frees the device buffers (waits for the work that reads them)."""
        ref, self._ref = self._ref, None
        if ref is not None and ref.value:
            try:
                ctx = get_context(self.dev)
                ctx.lib.b4d_phase_reference_destroy(ctx.handle, ref)
            except Exception:
                pass
        if _last_tracker.get(self.dev) is self:
            _last_tracker.pop(self.dev, None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def track(self, stack, *, subpixel: bool = True, return_device: bool = False):
        """(T, ny, nx) stack -> float64 table (T, 4) = dy, dx, peak, snr."""
        torch = require_cuda()
        check_stack(stack)
        T, ny, nx = stack.shape
        if (ny, nx) != (self.ny, self.nx):
            raise ValueError("frame shape differs from the tracker's reference frame shape")
        ctx = get_context(self.dev)
        out = torch.empty((T, 4), dtype=torch.float64, device=stack.device)
        ctx.check(ctx.lib.b4d_phase_track_ref(ctx.handle, self.handle, ptr(stack), T, ny, nx, int(bool(subpixel)), self.eps, 0,
                                              ptr(out)), "b4d_phase_track_ref")
        resolve_tracking(stack, out, self, subpixel=subpixel)
        return out if return_device else out.cpu().numpy()


# the tracker most recently created on each device: what stack_pipeline(want_tracking=True) uses when no tracker is passed
_last_tracker: dict[int, "PhaseTracker"] = {}


def check_stack(stack):
    """Device stacks are handed to the library as raw pointers: they must be float32, C-contiguous CUDA tensors."""
    torch = require_cuda()
    if not isinstance(stack, torch.Tensor) or stack.device.type != "cuda":
        raise TypeError("expected a CUDA tensor (use engine.as_stack for host data)")
    if stack.dtype != torch.float32:
        raise TypeError(f"device stacks must be float32, got {stack.dtype} (use engine.as_stack / cast_to_f32)")
    if not stack.is_contiguous():
        raise ValueError("device stacks must be C-contiguous (call .contiguous())")


def template_match(template, stack, *, ref_center_yx, subpixel: bool = True, eps: float = 1e-9, return_device: bool = False):
    """Normalised cross-correlation matching of one (h, w) template against every frame of a (T, ny, nx) stack, or of a
    (T, h, w) template stack frame by frame (b4d_template_match). ref_center_yx: centre of the template's reference position. -> (T, 4) = dy, dx, peak, snr."""
    torch = require_cuda()
    T, ny, nx = stack.shape
    check_fft_shape(ny, nx, generic_ok=True)
    tpl = as_stack(template, _dev(stack)).contiguous()
    per_frame = tpl.shape[0] != 1 or getattr(template, "ndim", 2) == 3
    if per_frame and tpl.shape[0] != T:
        raise ValueError("a template stack needs one template per frame")
    h, w = (int(v) for v in tpl.shape[1:])
    if h > ny or w > nx:
        raise ValueError(f"template shape {(h, w)} must fit inside image shape {(ny, nx)}")
    ctx = get_context(_dev(stack))
    out = torch.empty((T, 4), dtype=torch.float64, device=stack.device)
    ctx.check(ctx.lib.b4d_template_match(ctx.handle, ptr(tpl), int(per_frame), h, w, ptr(stack), T, ny, nx, float(ref_center_yx[0]),
                                         float(ref_center_yx[1]), int(bool(subpixel)), float(eps), ptr(out)),
              "b4d_template_match")
    return out if return_device else out.cpu().numpy()


def resolve_tracking(stack, track, tracker: "PhaseTracker | None" = None, *, subpixel: bool = True, flat_field_fn=None):
    """Frames whose fused median bracket missed carry snr = NaN (include/b4d.h, b4d_set_fused_median): redo them through the
    map-based exact path of the same tracker. `stack` holds the frames as the tracker saw them (raw when gain/dark were
    fused into the loaders: then `flat_field_fn` materialises the corrected frames). In place; returns the number of
    frames redone."""
    torch = require_cuda()
    bad = torch.nonzero(torch.isnan(track[:, 3])).flatten()
    if not bad.numel():
        return 0
    tracker = tracker if tracker is not None else _last_tracker.get(_dev(stack))
    if tracker is None:
        raise ValueError("resolve_tracking: no tracker")
    frames = stack[bad].contiguous()
    if flat_field_fn is not None:
        frames = flat_field_fn(frames)
    ctx = get_context(_dev(stack))
    T, ny, nx = frames.shape
    redo = torch.empty((T, 4), dtype=torch.float64, device=stack.device)
    ctx.check(ctx.lib.b4d_phase_track_ref(ctx.handle, tracker.handle, ptr(frames), T, ny, nx, int(bool(subpixel)),
                                          float(tracker.eps), 1, ptr(redo)), "b4d_phase_track_ref")
    track[bad] = redo
    return int(bad.numel())


def stack_pipeline(stack, *, gain=None, dark=None, saturation_value: float | None = 65535.0, eps: float = 1e-6,
                   psd_scale: float | None = None, subpixel: bool = True, track_eps: float = 1e-9,
                   want_reductions: bool = True, want_psd: bool = True, want_autocorr: bool = True,
                   want_grain: bool = True, want_tracking: bool = True, psd_out=None, ac_out=None,
                   tail_quantiles=None, tracker: "PhaseTracker | None" = None, want_spectral: bool = False):
    """The fused north-star pass over an HBM-resident stack (see b4d_stack_pipeline_ref in include/b4d.h).

    Tracking runs against `tracker` (its own reference spectrum); without one, against the PhaseTracker most recently
    created on the stack's device.
    tail_quantiles=(q_lo, q_hi) (fractions) adds the order statistics bracketing the two percentiles, collected in
    the reduction pass: "quantiles" (T, 4) float32 and "n_valid" (T,) int64 (-1 = unresolved frame, see
    resolve_tail_quantiles).
    want_spectral=True adds "spectral" (T, SP_NCOLS): the sums behind bandwidth() / spectral_entropy(), taken in the column
    pass of the same forward transform (power-of-two sides, needs the autocorrelation branch; f95 on square frames).
    Returns a dict with device tensors: reductions (T, FR_NCOLS), psd, autocorr, grain (T,4), tracking (T,4).
    """
    torch = require_cuda()
    check_stack(stack)
    T, ny, nx = stack.shape
    check_fft_shape(ny, nx, generic_ok=True)
    ctx = get_context(_dev(stack))
    dev = stack.device
    if want_tracking:
        tracker = tracker if tracker is not None else _last_tracker.get(_dev(stack))
        if tracker is None:
            raise ValueError("stack_pipeline: tracking needs a PhaseTracker")
        if (tracker.ny, tracker.nx) != (ny, nx):
            raise ValueError("frame shape differs from the tracker's reference frame shape")
    fr = torch.empty((T, FR_NCOLS), dtype=torch.float64, device=dev) if want_reductions else None
    if want_psd and psd_out is None:
        psd_out = torch.empty((T, ny, nx), dtype=torch.float32, device=dev)
    if want_autocorr and ac_out is None:
        ac_out = torch.empty((T, ny, nx), dtype=torch.float32, device=dev)
    grain = torch.empty((T, 4), dtype=torch.float64, device=dev) if want_grain else None
    track = torch.empty((T, 4), dtype=torch.float64, device=dev) if want_tracking else None
    sat = float("nan") if saturation_value is None else float(saturation_value)
    scale = (1.0 / (float(nx) * float(ny))) if psd_scale is None else float(psd_scale)
    quant = nvalid = None
    q_lo = q_hi = 0.0
    if tail_quantiles is not None:
        q_lo, q_hi = float(tail_quantiles[0]), float(tail_quantiles[1])
        quant = torch.empty((T, 4), dtype=torch.float32, device=dev)
        nvalid = torch.empty((T,), dtype=torch.int64, device=dev)
    spectral = torch.full((T, SP_NCOLS), float("nan"), dtype=torch.float64, device=dev) if want_spectral else None
    ctx.check(ctx.lib.b4d_stack_pipeline_ref(ctx.handle, tracker.handle if want_tracking else None, ptr(stack), T, ny, nx,
                                             ptr(gain), ptr(dark), sat, float(eps), scale, int(bool(subpixel)),
                                             float(tracker.eps if want_tracking else track_eps), q_lo, q_hi, ptr(fr),
                                             ptr(quant), ptr(nvalid), ptr(psd_out if want_psd else None),
                                             ptr(ac_out if want_autocorr else None), ptr(grain), ptr(track), ptr(spectral)),
              "b4d_stack_pipeline_ref")
    return {"reductions": fr, "psd": psd_out if want_psd else None, "autocorr": ac_out if want_autocorr else None,
            "grain": grain, "tracking": track, "quantiles": quant, "n_valid": nvalid, "spectral": spectral}


def frame_reductions_tails(stack, q_lo: float, q_hi: float, *, gain=None, dark=None,
                           saturation_value: float | None = 65535.0, eps: float = 1e-6):
    """Frame reductions + the tail order statistics of the same pass (b4d_frame_reductions_tails).
    Returns device tensors (table (T, FR_NCOLS) f64, quantiles (T, 4) f32, n_valid (T,) i64)."""
    torch = require_cuda()
    T, ny, nx = stack.shape
    ctx = get_context(_dev(stack))
    dev = stack.device
    fr = torch.empty((T, FR_NCOLS), dtype=torch.float64, device=dev)
    quant = torch.empty((T, 4), dtype=torch.float32, device=dev)
    nvalid = torch.empty((T,), dtype=torch.int64, device=dev)
    sat = float("nan") if saturation_value is None else float(saturation_value)
    ctx.check(ctx.lib.b4d_frame_reductions_tails(ctx.handle, ptr(stack), T, ny, nx, ptr(gain), ptr(dark), sat, float(eps),
                                                 float(q_lo), float(q_hi), ptr(fr), ptr(quant), ptr(nvalid)),
              "b4d_frame_reductions_tails")
    return fr, quant, nvalid


def resolve_tail_quantiles(stack, quant, nvalid, q_lo: float, q_hi: float):
    """Frames the fused tail collection flagged (n_valid == -1: the sample bracket missed, or the tails are not small)
    are redone by the exact stand-alone select; `stack` must hold the corrected pixels of those frames.  In place."""
    torch = require_cuda()
    bad = torch.nonzero(nvalid < 0).flatten()
    if bad.numel():
        q, nv = select_quantiles(stack[bad].contiguous(), [q_lo, q_hi], return_device=True)
        quant[bad] = q
        nvalid[bad] = nv
    return quant, nvalid
