"""
ctypes binding of libb4d.so (include/b4d.h) and the device plumbing around it.

PyTorch is used only as plumbing: device memory, streams, pinned host buffers and (elsewhere)
torch.distributed.  All arithmetic of the hot path happens inside libb4d.so; if the library is
missing or no CUDA device is present the product fails loudly -- there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B4D_LIB") or os.path.join(_HERE, "libb4d.so")   # B4D_LIB: A/B builds (build.py --variant)

FR_NCOLS = 13
FR = {"count": 0, "mean": 1, "m2": 2, "m3": 3, "m4": 4, "nzero": 5, "nsat": 6,
      "sgx2": 7, "sgy2": 8, "slap": 9, "slap2": 10, "npix": 11, "nnan": 12}
SP_NCOLS = 8
SP = {"total": 0, "fx2": 1, "fy2": 2, "p2": 3, "all": 4, "plogp": 5, "f95": 6}
FFT_MIN, FFT_MAX = 128, 2048


class B4DError(RuntimeError):
    pass


class B4DUnsupported(NotImplementedError):
    pass


_lib = None
_lib_lock = threading.Lock()

_vp, _i64, _i32, _f32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/b4d.h one to one
_SIGNATURES = {
    "b4d_create": [_i32, C.POINTER(_vp)],
    "b4d_destroy": [_vp],
    "b4d_set_stream": [_vp, _vp],
    "b4d_synchronize": [_vp],
    "b4d_last_error": [_vp],
    "b4d_version": [],
    "b4d_launch_count": [_vp],
    "b4d_device_sm_count": [_vp],
    "b4d_profile_begin": [_vp],
    "b4d_profile_end": [_vp, _vp, _vp],
    "b4d_profile_class_name": [_i32],
    "b4d_set_batch_frames": [_vp, _i64],
    "b4d_set_schedule": [_vp, _i32, _i32, _i32, _i32, _i32],
    "b4d_set_pairing": [_vp, _i32],
    "b4d_set_fused_median": [_vp, _i32],
    "b4d_malloc": [_vp, C.c_size_t, C.POINTER(_vp)],
    "b4d_free": [_vp, _vp],
    "b4d_memcpy_h2d": [_vp, _vp, _vp, C.c_size_t],
    "b4d_memcpy_d2h": [_vp, _vp, _vp, C.c_size_t],
    "b4d_cast_to_f32": [_vp, _vp, _i32, _vp, _i64],
    "b4d_inflate_caps": [_vp, C.POINTER(_i32), C.POINTER(_i64)],
    "b4d_inflate_batch": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64],
    "b4d_unchunk_to_f32": [_vp, _vp, _i32, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "b4d_frame_reductions": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _f64, _f64, _vp],
    "b4d_select_ranks": [_vp, _vp, _i64, _i64, _vp, _i32, _i32, _vp, _vp],
    "b4d_flat_field": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _f32, _i32, _vp],
    "b4d_bad_pixel_repair": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _f32],
    "b4d_flat_gain": [_vp, _vp, _vp, _i32, _i32, _f32, _f32, _vp],
    "b4d_sub": [_vp, _vp, _vp, _i64, _vp],
    "b4d_temporal_accumulate": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp],
    "b4d_temporal_pilot": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp],
    "b4d_temporal_finalize": [_vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "b4d_fft2d": [_vp, _vp, _i64, _i32, _i32, _vp],
    "b4d_ifft2d": [_vp, _vp, _i64, _i32, _i32, _vp],
    "b4d_psd2d": [_vp, _vp, _i64, _i32, _i32, _f32, _i32, _i32, _vp, _vp],
    "b4d_autocorr2d": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _f64, _vp],
    "b4d_xcorr2d": [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp],
    "b4d_template_match": [_vp, _vp, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _f64, _f64, _i32, _f64, _vp],
    "b4d_phase_set_reference": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f64],
    "b4d_phase_track": [_vp, _vp, _i64, _i32, _i32, _i32, _f64, _vp],
    "b4d_phase_reference_create": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f64, C.POINTER(_vp)],
    "b4d_phase_reference_destroy": [_vp, _vp],
    "b4d_phase_track_ref": [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _f64, _i32, _vp],
    "b4d_stack_pipeline": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _f64, _f64, _f32, _i32, _f64, _f64, _f64,
                           _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "b4d_stack_pipeline_ref": [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _f64, _f64, _f32, _i32, _f64, _f64, _f64,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "b4d_frame_reductions_tails": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _f64, _f64, _f64, _f64, _vp, _vp, _vp],
}
_RESTYPES = {"b4d_profile_class_name": C.c_char_p, "b4d_last_error": C.c_char_p, "b4d_version": C.c_char_p, "b4d_launch_count": _i64}


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libb4d.so and attach the prototypes.  Raises B4DError when the library is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise B4DError(
                f"{p} not found: build it with `python -m barc4dip_b200.build` "
                "(nvcc, sm_100a). barc4dip_b200 has no CPU fallback.")
        lib = C.CDLL(p)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library drift
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _i32)
        if path is None:
            _lib = lib
        return lib


# entry points that do no stream-ordered work: the bound library does not re-install the stream for them
_NO_STREAM = {"b4d_inflate_caps", "b4d_create", "b4d_destroy", "b4d_set_stream", "b4d_last_error", "b4d_version", "b4d_launch_count",
              "b4d_device_sm_count", "b4d_profile_class_name", "b4d_set_batch_frames", "b4d_set_schedule", "b4d_set_pairing",
              "b4d_set_fused_median"}


class _BoundLib:
    """The library as one context sees it: every compute call installs the calling thread's current torch stream and
    runs under the context's (re-entrant) Python lock, so that `set stream` + `launch` is one step even when several
    threads share the context (the reference drives per-frame calls from joblib threads, metrics/speckles.py:323)."""

    def __init__(self, ctx: "Context", raw):
        self._ctx, self._raw, self._fns = ctx, raw, {}

    def __getattr__(self, name):
        fn = self._fns.get(name)
        if fn is None:
            raw_fn = getattr(self._raw, name)
            if name in _NO_STREAM:
                fn = raw_fn
            else:
                ctx, raw = self._ctx, self._raw

                def fn(*args, _f=raw_fn):
                    import torch
                    with ctx.lock:
                        if ctx.pin_stream is None:
                            raw.b4d_set_stream(ctx.handle, _vp(torch.cuda.current_stream(ctx.device).cuda_stream))
                        return _f(*args)
            self._fns[name] = fn
        return fn


class Context:
    """One libb4d context = (device, stream, scratch).  Calls on one context are serialised."""

    def __init__(self, device: int):
        self.raw = load_library()
        self.lib = _BoundLib(self, self.raw)
        self.lock = threading.RLock()
        self.pin_stream = None      # set by use_stream(): launches go to that stream whatever torch's current stream is
        self.device = int(device)
        h = _vp()
        rc = self.raw.b4d_create(self.device, C.byref(h))
        if rc != 0 or not h.value:
            raise B4DError(f"b4d_create(device={device}) failed with status {rc} "
                           "(is a CUDA device visible?)")
        self.handle = h

    def check(self, rc: int, what: str = ""):
        if rc == 0:
            return
        msg = self.lib.b4d_last_error(self.handle)
        text = msg.decode() if msg else ""
        if rc == -3:
            raise B4DUnsupported(f"{what}: {text}")
        if rc == -1:
            raise ValueError(f"{what}: {text}")
        raise B4DError(f"{what} failed (status {rc}): {text}")

    def use_current_stream(self):
        """Kept for callers of round 1: every compute call now installs the calling thread's current stream itself."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        with self.lock:
            self.check(self.raw.b4d_set_stream(self.handle, _vp(s)), "b4d_set_stream")

    def profile_begin(self):
        self.check(self.lib.b4d_profile_begin(self.handle), "b4d_profile_begin")

    def profile_end(self) -> dict:
        """{class name: (milliseconds, launches)} of the kernels launched since profile_begin()."""
        n = 16   # B4D_PROF_NCLASS
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self.check(self.lib.b4d_profile_end(self.handle, ms, cnt), "b4d_profile_end")
        return {self.lib.b4d_profile_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def set_batch_frames(self, frames: int):
        self.check(self.lib.b4d_set_batch_frames(self.handle, int(frames)), "b4d_set_batch_frames")

    def set_schedule(self, sub_frames: int = -1, lanes: int = -1, ring_slots: int = -1, keep: int = -1, use_graphs: int = -1):
        """Frame-pipelined schedule of the fused stack pipeline (include/b4d.h, b4d_set_schedule); -1 = default."""
        self.check(self.lib.b4d_set_schedule(self.handle, int(sub_frames), int(lanes), int(ring_slots), int(keep),
                                             int(use_graphs)), "b4d_set_schedule")

    def set_pairing(self, pair_frames: int = -1):
        self.check(self.lib.b4d_set_pairing(self.handle, int(pair_frames)), "b4d_set_pairing")

    def set_fused_median(self, on: bool):
        self.check(self.lib.b4d_set_fused_median(self.handle, int(bool(on))), "b4d_set_fused_median")

    @property
    def launches(self) -> int:
        return int(self.lib.b4d_launch_count(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.raw.b4d_destroy(self.handle)
                self.handle = _vp()
        except Exception:
            pass


_contexts: dict[int, Context] = {}
_ctx_lock = threading.Lock()


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise B4DError("barc4dip_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                       "there is no CPU fallback on this path.")
    return torch


def default_device() -> int:
    """Device index: $B4D_DEVICE, else the current torch device (under torchrun: what torch.cuda.set_device(LOCAL_RANK) chose)."""
    torch = require_cuda()
    env = os.environ.get("B4D_DEVICE")
    if env is not None:
        return int(env)
    return torch.cuda.current_device()


def get_context(device: int | None = None) -> Context:
    require_cuda()
    dev = default_device() if device is None else int(device)
    with _ctx_lock:
        ctx = _contexts.get(dev)
        if ctx is None:
            ctx = Context(dev)
            _contexts[dev] = ctx
    return ctx


def ptr(t) -> _vp:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return _vp(0)
    return _vp(t.data_ptr())


# integer types b4d_cast_to_f32 widens on the device (enum b4d_dtype in include/b4d.h)
NATIVE_INT_CODES = {"uint8": 0, "uint16": 1, "int16": 2, "int32": 3, "uint32": 4}


def native_int_code(dtype) -> int | None:
    """b4d_dtype code of a numpy / torch integer dtype the device cast accepts, else None."""
    name = str(dtype).replace("torch.", "")
    return NATIVE_INT_CODES.get(name)


def cast_to_f32(raw_bytes, code: int, out):
    """raw_bytes: CUDA uint8 tensor holding `out.numel()` elements of integer type `code`; out: float32 CUDA tensor."""
    ctx = get_context(out.device.index or 0)
    ctx.check(ctx.lib.b4d_cast_to_f32(ctx.handle, ptr(raw_bytes), int(code), ptr(out), int(out.numel())), "b4d_cast_to_f32")
    return out


def inflate_caps(device: int | None = None) -> tuple[int, int]:
    """(algorithm mask, largest stream in bytes) of the device's hardware decompression engine; mask bit 0 = deflate.
    (0, 0) where the device or the driver has none."""
    ctx = get_context(device)
    mask, mx = _i32(0), _i64(0)
    ctx.check(ctx.lib.b4d_inflate_caps(ctx.handle, C.byref(mask), C.byref(mx)), "b4d_inflate_caps")
    return int(mask.value), int(mx.value)


UNCHUNK_CODES = dict(NATIVE_INT_CODES, float32=5)


def as_device_f32(a, device: int | None = None):
    """numpy array / torch tensor -> contiguous float32 CUDA tensor (H2D copy when given host data). Integer detector
    types (uint8/uint16/int16/int32/uint32) cross PCIe in their own width and are widened on the device."""
    torch = require_cuda()
    dev = default_device() if device is None else int(device)
    if (isinstance(a, torch.Tensor) and a.is_complex()) or (not isinstance(a, torch.Tensor) and np.iscomplexobj(a)):
        raise TypeError("complex input is not supported on this path (real frames only)")
    if not isinstance(a, torch.Tensor):
        arr = np.asarray(a)
        code = native_int_code(arr.dtype)
        if code is not None and arr.size >= 4096:
            src = np.ascontiguousarray(arr)
            raw = torch.from_numpy(src.reshape(-1).view(np.uint8)).to(f"cuda:{dev}")
            return cast_to_f32(raw, code, torch.empty(src.shape, dtype=torch.float32, device=raw.device))
    if isinstance(a, torch.Tensor):
        t = a
        if t.device.type != "cuda":
            t = t.to(f"cuda:{dev}", non_blocking=True)
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        return t.contiguous()
    arr = np.asarray(a)
    if arr.dtype == np.float32:
        src = np.ascontiguousarray(arr)
    else:
        src = np.ascontiguousarray(arr, dtype=np.float32)
    return torch.from_numpy(src).to(f"cuda:{dev}", non_blocking=False)
