"""
The public stack-analysis call: host stack in, host results out, frames resident in HBM in between.

    analyzer = StackAnalyzer((ny, nx), reference=frame0)        # reference frame for the tracker
    res = analyzer.run(stack)                                    # numpy (T, ny, nx) or CUDA tensor

One fused pass per chunk of frames (b4d_stack_pipeline) produces, per frame,
  * the single-pass reductions (distribution moments, Tenengrad, Laplacian variance, visibility),
  * the exact 0.05 / 99.95 percentile contrast (amplitude),
  * the PSD map (psd2d) and the peak-normalised autocorrelation map (autocorr2d) with the grain widths,
  * the phase-correlation displacement against the reference frame (dy, dx, peak, snr).
Host input is staged through pinned memory in chunks; the H2D copy of chunk c+1 and the D2H copy of chunk
c-1 overlap the kernels of chunk c on separate CUDA streams.

This is the reference's per-frame Python loop (metrics/speckles.py:300-325, :347-386) collapsed into batched
kernels; results use the reference's names and conventions.
"""

from __future__ import annotations

import numpy as np

from . import engine, stack as blocks
from ._lib import FR_NCOLS, cast_to_f32, get_context, native_int_code, require_cuda


class StackAnalyzer:
    def __init__(self, frame_shape, *, reference=None, device: int | None = None, chunk_frames: int = 8,
                 want_maps: bool = True, want_contrast: bool = True, saturation_value: float | None = 65535.0,
                 eps: float = 1e-6, subpixel: bool = True, flats=None, darks=None, scale: str = "flat_median",
                 flat_eps: float | None = None):
        torch = require_cuda()
        from ._lib import default_device
        self.dev = default_device() if device is None else int(device)
        self.ny, self.nx = int(frame_shape[0]), int(frame_shape[1])
        engine.check_fft_shape(self.ny, self.nx, generic_ok=True)
        self.chunk = int(chunk_frames)
        self.want_maps, self.want_contrast = bool(want_maps), bool(want_contrast)
        self.sat, self.eps, self.subpixel = saturation_value, float(eps), bool(subpixel)
        self.device = torch.device(f"cuda:{self.dev}")
        self.gain = self.dark = self.flat = None
        self._ff = None
        if flats is not None:
            from .preprocessing.normalize import _collapse, resolve_flat_field
            flat, dark = _collapse(flats, self.dev), _collapse(darks, self.dev)
            e, s, apply_scale = resolve_flat_field(flat, dark, scale=scale, eps=flat_eps)
            self.flat, self.dark = flat, (dark if dark is not None else torch.zeros_like(flat))
            self._ff = dict(eps=e, scale_value=s, apply_scale=apply_scale)
            # per-pixel multiplier used by the fused loaders: (raw - dark) * gain
            self.gain = engine.flat_gain(flat, dark, eps=e, scale_value=s if apply_scale else 1.0)
        self.tracker = None
        if reference is not None:
            self.set_reference(reference)
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(3)]   # h2d, compute, d2h
        self._stage = None

    # the reference frame is corrected like every other frame when a flat field is installed
    def set_reference(self, frame, *, slices_yx=None):
        torch = require_cuda()
        ref = engine.as_stack(frame, self.dev)
        if self.gain is not None:
            ref = engine.flat_field(ref, self.flat, self.dark, **self._ff)
        ref = ref[0]
        if slices_yx is None:
            y0 = x0 = 0
            tpl = ref
        else:
            sy, sx = slices_yx
            y0, x0 = int(sy.start), int(sx.start)
            tpl = ref[sy, sx].contiguous()
        self.tracker = engine.PhaseTracker(tpl, (self.ny, self.nx), y0=y0, x0=x0, device=self.dev)

    def _buffers(self, pinned_out: bool):
        torch = require_cuda()
        if self._stage is None:
            c, ny, nx = self.chunk, self.ny, self.nx
            dev = self.device
            self._stage = {
                "in": [torch.empty((c, ny, nx), dtype=torch.float32, device=dev) for _ in range(2)],
                "psd": [torch.empty((c, ny, nx), dtype=torch.float32, device=dev) for _ in range(2)] if self.want_maps else None,
                "ac": [torch.empty((c, ny, nx), dtype=torch.float32, device=dev) for _ in range(2)] if self.want_maps else None,
            }
        return self._stage

    def run_device(self, dev_stack, *, psd_out=None, ac_out=None, resolve_tails: bool = True) -> dict:
        """Analyse an HBM-resident (T, ny, nx) float32 stack; everything stays on the device."""
        q_lo, q_hi = 0.05 / 100.0, 99.95 / 100.0      # amplitude(): percentile_minmax_range defaults
        # grain widths need square frames (the reference pads to square, metrics/speckles.py:530): the fused pass reports
        # them for square frames only, the others go through stack.grain_block below
        square = self.ny == self.nx
        res = engine.stack_pipeline(dev_stack, gain=self.gain, dark=self.dark, saturation_value=self.sat, eps=self.eps,
                                    subpixel=self.subpixel, want_psd=self.want_maps or psd_out is not None,
                                    want_autocorr=square or self.want_maps or ac_out is not None, want_grain=square,
                                    want_tracking=self.tracker is not None, tracker=self.tracker,
                                    psd_out=psd_out, ac_out=ac_out,
                                    tail_quantiles=(q_lo, q_hi) if self.want_contrast else None)
        if res["grain"] is None:
            # non-square frames: the reference pads to a square with the frame's mean (geometry/masks.py:11-56) before the
            # autocorrelation; stack.grain_block does that (the padded side is rarely a power of two: chirp-z path)
            torch = require_cuda()
            src = dev_stack
            if self.gain is not None:
                src = engine.flat_field(dev_stack, self.flat, self.dark, **self._ff)
            g = blocks.grain_block(src, table=res["reductions"].cpu().numpy())
            res["grain"] = torch.from_numpy(np.stack([g["lx"], g["ly"], g["leq"], g["r"]], axis=1)).to(dev_stack.device)
        if resolve_tails and res["tracking"] is not None:
            # same contract for the tracker: snr = NaN marks a frame whose fused median bracket missed
            ff = (lambda fr: engine.flat_field(fr, self.flat, self.dark, **self._ff)) if self.gain is not None else None
            engine.resolve_tracking(dev_stack, res["tracking"], self.tracker, subpixel=self.subpixel, flat_field_fn=ff)
        if self.want_contrast and resolve_tails:
            # the tails are collected inside the reduction pass; a frame the sample bracket missed (n_valid == -1) is
            # redone by the stand-alone exact select (host sync: one tiny D2H per chunk)
            nv = res["n_valid"]
            if bool((nv < 0).any()):
                src = dev_stack
                if self.gain is not None:  # the stand-alone select needs the corrected pixels materialised
                    bad = (nv < 0).nonzero().flatten()
                    src = dev_stack.clone()
                    src[bad] = engine.flat_field(dev_stack[bad].contiguous(), self.flat, self.dark, **self._ff)
                engine.resolve_tail_quantiles(src, res["quantiles"], nv, q_lo, q_hi)
        return res

    def run(self, stack, *, keep_maps_on_device: bool = False, reuse_host_buffers: bool = False) -> dict:
        """(T, ny, nx) numpy (or CUDA tensor) -> dict of numpy results (maps as float32 arrays, or device tensors).

        reuse_host_buffers=True returns the PSD / autocorrelation maps as views of pinned buffers the analyzer keeps and
        overwrites on its next run() of the same stack length (no 34 MB-per-frame allocation per call); by default every
        call returns maps of its own."""
        torch = require_cuda()
        is_host = not (isinstance(stack, torch.Tensor) and stack.device.type == "cuda")
        T = int(stack.shape[0])
        if tuple(stack.shape[1:]) != (self.ny, self.nx):
            raise ValueError("frame shape differs from the analyzer's")
        if not is_host:
            # a device stack is handed over as a raw pointer: float32, contiguous, on this analyzer's device
            if stack.device != self.device:
                raise ValueError(f"the stack lives on {stack.device}, the analyzer on {self.device}")
            if stack.dtype != torch.float32 or not stack.is_contiguous():
                stack = stack.to(torch.float32).contiguous()
        ny, nx, c = self.ny, self.nx, self.chunk
        h2d, comp, d2h = self._streams
        # everything the caller (and this object's constructor / set_reference) queued on the current stream -- the
        # stack itself when it is a device tensor, the gain map, the reference spectrum -- precedes the private streams
        cur = torch.cuda.current_stream(self.device)
        for s_ in self._streams:
            s_.wait_stream(cur)
        st = self._buffers(True)
        code = None
        if is_host:
            # float32 stacks are staged as they are; integer detector types travel in their own width and are widened
            # on the device (b4d_cast_to_f32); anything else is converted to float32 on the host first
            code = native_int_code(stack.dtype)
            if isinstance(stack, torch.Tensor):
                src = stack.contiguous() if (code is not None or stack.dtype == torch.float32) else stack.to(torch.float32)
            elif code is not None:
                src = torch.from_numpy(np.ascontiguousarray(stack).reshape(T, -1).view(np.uint8))
            else:
                src = torch.from_numpy(np.ascontiguousarray(stack, dtype=np.float32))
            if code is not None:
                if isinstance(stack, torch.Tensor):
                    src = src.reshape(T, -1).view(torch.uint8)
                if st.get("raw") is None or st["raw"][0].shape[1] != src.shape[1]:
                    st["raw"] = [torch.empty((c, src.shape[1]), dtype=torch.uint8, device=self.device) for _ in range(2)]
        fr_all = torch.empty((T, FR_NCOLS), dtype=torch.float64, device=self.device)
        grain_all = torch.empty((T, 4), dtype=torch.float64, device=self.device)
        track_all = torch.empty((T, 4), dtype=torch.float64, device=self.device) if self.tracker is not None else None
        q_all = torch.empty((T, 4), dtype=torch.float32, device=self.device) if self.want_contrast else None
        nv_all = torch.empty((T,), dtype=torch.int64, device=self.device) if self.want_contrast else None
        maps_host = None
        if self.want_maps and not keep_maps_on_device:
            # pinning 34 MB per frame is not free: with reuse_host_buffers the pinned result buffers are cached per stack
            # length (and the returned maps alias them until the next run)
            cache = getattr(self, "_host_maps", None) if reuse_host_buffers else None
            if cache is None or cache["psd"].shape[0] != T:
                cache = {k: torch.empty((T, ny, nx), dtype=torch.float32, pin_memory=True) for k in ("psd", "ac")}
                if reuse_host_buffers:
                    self._host_maps = cache
            maps_host = cache
        maps_dev = {k: torch.empty((T, ny, nx), dtype=torch.float32, device=self.device) for k in ("psd", "ac")} \
            if (self.want_maps and keep_maps_on_device) else None

        ctx = get_context(self.dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        n_chunks = (T + c - 1) // c
        for i in range(n_chunks):
            a, b = i * c, min(T, (i + 1) * c)
            n = b - a
            s = i & 1
            if is_host:
                with torch.cuda.stream(h2d):
                    if i >= 2:
                        h2d.wait_event(ev_done[s])          # the kernels that read this staging buffer have finished
                    (st["raw"] if code is not None else st["in"])[s][:n].copy_(src[a:b], non_blocking=True)
                    ev_in[s].record(h2d)
                frames = st["in"][s][:n]
            else:
                frames = stack[a:b]
            with torch.cuda.stream(comp):
                if is_host:
                    comp.wait_event(ev_in[s])
                if self.want_maps and not keep_maps_on_device and i >= 2:
                    comp.wait_event(ev_out[s])              # the D2H that drained this map buffer has finished
                if code is not None:
                    cast_to_f32(st["raw"][s][:n], code, frames)
                if self.want_maps:
                    po = maps_dev["psd"][a:b] if keep_maps_on_device else st["psd"][s][:n]
                    ao = maps_dev["ac"][a:b] if keep_maps_on_device else st["ac"][s][:n]
                else:
                    po = ao = None
                res = self.run_device(frames, psd_out=po, ac_out=ao, resolve_tails=False)
                fr_all[a:b].copy_(res["reductions"])
                grain_all[a:b].copy_(res["grain"])
                if track_all is not None:
                    track_all[a:b].copy_(res["tracking"])
                if q_all is not None:
                    q_all[a:b].copy_(res["quantiles"])
                    nv_all[a:b].copy_(res["n_valid"])
                ev_done[s].record(comp)
            if maps_host is not None:
                with torch.cuda.stream(d2h):
                    d2h.wait_event(ev_done[s])
                    maps_host["psd"][a:b].copy_(st["psd"][s][:n], non_blocking=True)
                    maps_host["ac"][a:b].copy_(st["ac"][s][:n], non_blocking=True)
                    ev_out[s].record(d2h)
        for s_ in self._streams:
            s_.synchronize()
        if track_all is not None and bool(torch.isnan(track_all[:, 3]).any()):
            # frames whose fused median bracket missed: redo them through the map-based tracker
            bad = torch.isnan(track_all[:, 3]).nonzero().flatten()
            frames = self._frames_of(stack, bad)
            sub = track_all[bad].clone()
            ff = (lambda fr: engine.flat_field(fr, self.flat, self.dark, **self._ff)) if self.gain is not None else None
            engine.resolve_tracking(frames, sub, self.tracker, subpixel=self.subpixel, flat_field_fn=ff)
            track_all[bad] = sub
        if nv_all is not None and bool((nv_all < 0).any()):
            # frames whose tails the fused collection did not resolve: exact stand-alone select on those frames only
            bad = (nv_all < 0).nonzero().flatten()
            frames = self._frames_of(stack, bad)
            if self.gain is not None:
                frames = engine.flat_field(frames, self.flat, self.dark, **self._ff)
            q, nv = engine.select_quantiles(frames, [0.05 / 100.0, 99.95 / 100.0], return_device=True)
            q_all[bad] = q
            nv_all[bad] = nv

        fr = fr_all.cpu().numpy()
        out = {"stats": blocks.moments_block(fr, self.sat), "gradient": blocks.gradient_block(fr),
               "laplacian": blocks.laplacian_block(fr), "table": fr}
        g = grain_all.cpu().numpy()
        out["grain"] = {"lx": g[:, 0], "ly": g[:, 1], "leq": g[:, 2], "r": g[:, 3]}
        with np.errstate(invalid="ignore", divide="ignore"):
            vis = np.sqrt(fr[:, 2]) / fr[:, 1]
        amp = {"visibility": vis}
        if q_all is not None:
            q, nv = q_all.cpu().numpy(), nv_all.cpu().numpy()
            con = np.empty(T)
            for t in range(T):
                lo = engine.quantile_from_bracket(q[t, 0], q[t, 1], int(nv[t]), 0.05 / 100.0)
                hi = engine.quantile_from_bracket(q[t, 2], q[t, 3], int(nv[t]), 99.95 / 100.0)
                con[t] = (hi - lo) / (hi + lo) if (hi + lo) > 0 else np.nan
            amp["contrast"] = con
        out["amplitude"] = amp
        if track_all is not None:
            tr = track_all.cpu().numpy()
            out["tracking"] = {"dy": tr[:, 0], "dx": tr[:, 1], "peak": tr[:, 2], "snr": tr[:, 3]}
        if maps_host is not None:
            out["psd"], out["autocorr"] = maps_host["psd"].numpy(), maps_host["ac"].numpy()
        elif maps_dev is not None:
            out["psd"], out["autocorr"] = maps_dev["psd"], maps_dev["ac"]
        return out

    def _frames_of(self, stack, idx):
        """Frames `idx` (1-D index tensor) of the caller's stack as a float32 device stack (fallback paths only)."""
        torch = require_cuda()
        if isinstance(stack, torch.Tensor):
            return stack[idx.to(stack.device)].to(self.device).to(torch.float32).contiguous()
        return engine.as_stack(np.asarray(stack)[idx.cpu().numpy()], self.dev)

    def bytes_per_frame(self) -> tuple[int, int]:
        """(host->device, device->host) bytes per frame of run() with host input."""
        npix = self.ny * self.nx
        d2h = (FR_NCOLS + 4 + (4 if self.tracker is not None else 0)) * 8 + (4 * 4 + 8 if self.want_contrast else 0)
        if self.want_maps:
            d2h += 2 * npix * 4
        return npix * 4, d2h


def analyze_stack(stack, *, reference=None, **kw) -> dict:
    """Convenience wrapper: analyse a (T, ny, nx) stack against `reference` (default: its first frame)."""
    a = StackAnalyzer(stack.shape[1:], reference=stack[0] if reference is None else reference, **kw)
    return a.run(stack)
