#!/usr/bin/env python
"""
Benchmark of the barc4dip stack-analysis hot path on B200 (and of the reference's CPU path).

    python bench.py --gpus N --steps K --warmup W                 # our arm (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): stack frames/s at 2048^2 for the fused FFT-PSD + tracking + metrics pipeline.
A step = one pass of the hot path over one batch of synthetic frames per GPU:
    frame reductions (moments, Tenengrad, Laplacian variance, visibility) + percentile contrast
    + psd2d map + autocorr2d map with grain widths + phase-correlation tracking against a broadcast
    reference frame + accumulation of the per-pixel temporal power sums.
One timed run = one stack: broadcast of the reference frame and of the temporal shift plane (NCCL), K steps, then the
all-reduce of the four float64 power-sum planes (NCCL) and the temporal finalize -- all inside the timed region.
`value` is measured with the batch resident in HBM; `e2e` goes through the public API (StackAnalyzer.run)
from pinned host memory, host<->device copies of inputs and of every result inside the timed region.
One JSON line is printed by rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stack frames/sec at 2048^2 (FFT-PSD+tracking+metrics)"
UNIT = "frames/s"
MB = 1e6
ALGO_BYTES_PER_FRAME = {          # compulsory HBM traffic per frame, SURVEY.md 8(d) / DESIGN.md section 4
    "step": 3 * 2048 * 2048 * 4,            # read frame + write PSD + write autocorrelation = 50.33 MB
    "frame_reduce": 2048 * 2048 * 4,        # one streaming read
    "rows_fwd": 2048 * 2048 * 4,            # reads the frame, writes an L2-resident intermediate
    "cols": 2048 * 2048 * 4,                # writes the PSD map; intermediates and reference spectrum in L2
    "rows_inv": 2048 * 2048 * 4,            # writes the autocorrelation map
    "select_hist": 2048 * 2048 * 4,         # one pass over a frame-sized map
}
FFT_FLOPS_2048 = 5.0 * 2048 * 11            # 5 N log2 N per complex transform of 2048 points


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--frames", type=int, default=512, help="frames per step per GPU (resident in HBM)")
    p.add_argument("--e2e-frames", type=int, default=128, help="frames of the pinned host stack of the end-to-end legs")
    p.add_argument("--size", type=int, default=2048)
    p.add_argument("--batch", type=int, default=0, help="frames per internal FFT batch (0 = automatic)")
    p.add_argument("--e2e-steps", type=int, default=5)
    p.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = automatic)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--prewarm", type=float, default=1.5, help="seconds of untimed load before the warm-up steps (clock ramp)")
    return p.parse_args()


# --------------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------------

def make_shifts(n_frames: int, seed: int) -> np.ndarray:
    """2-D random walk, step sigma = 0.3 px, clipped to +-20 px; frame 0 unshifted. Every 8th frame is
    rounded to whole pixels (np.roll frames: known answers)."""
    rng = np.random.default_rng(seed)
    steps = rng.normal(0.0, 0.3, size=(n_frames, 2))
    steps[0] = 0.0
    shifts = np.clip(np.cumsum(steps, axis=0), -20.0, 20.0)
    shifts[::8] = np.round(shifts[::8] * 4.0)
    shifts[0] = 0.0
    return shifts


def cpu_stack(base: np.ndarray, shifts: np.ndarray, noise_seed: int) -> np.ndarray:
    """Host stack by the same recipe as the device one: Fourier-shifted copies of `base` + 1 % Gaussian noise on every
    frame (frame 0 included)."""
    from barc4dip_b200 import synth
    rng = np.random.default_rng(noise_seed)
    sigma = 0.01 * float(base.mean())
    out = np.empty((len(shifts),) + base.shape, np.float32)
    for t, (dy, dx) in enumerate(shifts):
        out[t] = synth.fourier_shift(base, float(dy), float(dx)) + rng.normal(0, sigma, base.shape).astype(np.float32)
    return out


def device_stack(base: np.ndarray, shifts: np.ndarray, noise_seed: int, dev):
    """The C4 stack built on the device: frame t = IFFT(FFT(base) * ramp(shift_t)) + noise, through the library's own
    fft2d / ifft2d (whole-pixel shifts are exact rolls)."""
    import torch
    from barc4dip_b200 import engine
    n0, n1 = base.shape
    g = torch.Generator(device=dev)
    g.manual_seed(noise_seed)
    base_d = torch.from_numpy(base).to(dev)
    spec = engine.fft2d(base_d[None])[0]                                # shifted spectrum, zero frequency at [n/2, n/2]
    ky = ((torch.arange(n0, device=dev, dtype=torch.float64) - n0 // 2) / n0)[:, None]
    kx = ((torch.arange(n1, device=dev, dtype=torch.float64) - n1 // 2) / n1)[None, :]
    sigma = 0.01 * float(base.mean())
    out = torch.empty((len(shifts), n0, n1), dtype=torch.float32, device=dev)
    for t, (dy, dx) in enumerate(shifts):
        if float(dy).is_integer() and float(dx).is_integer():
            fr = torch.roll(base_d, (int(dy), int(dx)), dims=(0, 1))
        else:
            ph = -2.0 * np.pi * (ky * float(dy) + kx * float(dx))
            ramp = torch.complex(torch.cos(ph), torch.sin(ph)).to(torch.complex64)
            fr = engine.ifft2d((spec * ramp)[None])[0].real
        out[t] = fr + sigma * torch.randn((n0, n1), generator=g, device=dev)
    return out


# --------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's per-frame path, joblib threads
# --------------------------------------------------------------------------------------------

def cpu_frame_work(frame: np.ndarray, ref: np.ndarray):
    from oracle import ref_numpy as orc
    n0, n1 = frame.shape
    m = orc.distribution_moments(frame)
    tg = orc.tenengrad(frame)
    lv = orc.laplacian_variance(frame)
    am = orc.amplitude(frame)
    P, _, _ = orc.psd2d(frame)
    g = orc.grain(frame)
    tr = orc.phase_correlation(ref, frame, slices_yx=(slice(0, n0), slice(0, n1)))
    return m["mean"], tg["tenengrad"], lv, am["contrast"], float(P[0, 0]), g["lx"], tr[0], tr[1]


def cpu_temporal_rows(stack: np.ndarray, shift: np.ndarray, r0: int, r1: int):
    """Per-pixel shifted power sums (row T of SURVEY 8(a): no reference function; numpy float64) of rows [r0, r1)."""
    d = stack[:, r0:r1, :].astype(np.float64) - shift[None, r0:r1, :]
    d2 = d * d
    return d.sum(axis=0), d2.sum(axis=0), (d2 * d).sum(axis=0), (d2 * d2).sum(axis=0)


def cpu_pipeline(stack: np.ndarray, ref: np.ndarray, n_jobs: int):
    """The reference drives frames with joblib threads (metrics/speckles.py:323, metrics/sharpness.py:361)."""
    from joblib import Parallel, delayed
    out = Parallel(n_jobs=n_jobs, prefer="threads")(delayed(cpu_frame_work)(stack[t], ref) for t in range(stack.shape[0]))
    shift = stack[: min(16, stack.shape[0])].mean(axis=0, dtype=np.float64)
    rows = np.linspace(0, stack.shape[1], n_jobs + 1).astype(int)
    Parallel(n_jobs=n_jobs, prefer="threads")(delayed(cpu_temporal_rows)(stack, shift, int(a), int(b)) for a, b in zip(rows[:-1], rows[1:]) if b > a)
    return out


def time_cpu(size: int, frames: int, steps: int, warmup: int):
    from barc4dip_b200 import synth
    cores = os.cpu_count() or 1
    base = synth.speckle_frame(size, grain=6.0, seed=0)
    stack = cpu_stack(base, make_shifts(frames, 2), 3)
    ref = stack[0]
    for _ in range(warmup):
        cpu_pipeline(stack[: max(1, min(frames, cores))], ref, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pipeline(stack, ref, cores)
    dt = time.perf_counter() - t0
    return frames * steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: about 400 frames in total (~2 minutes at the 3-4 frames/s this path reaches on 16 cores), so
    # that any --steps K finishes within a few minutes; 2..32 frames per step
    frames = args.cpu_frames or max(2, min(32, 400 // max(1, args.steps)))
    fps, sec_per_step, cores = time_cpu(args.size, frames, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"fused stack pipeline (moments+tenengrad+laplacian+amplitude+psd2d+grain/autocorr2d+phase_correlation"
                               f"+per-pixel temporal power sums), "
                               f"{args.size}x{args.size} float32 frames",
                   "stack": "SURVEY 8(d) C4: one speckle Fourier-shifted along a 2-D random walk (sigma 0.3 px, +-20 px) + 1 % noise per frame",
                   "frame": [args.size, args.size], "frames_per_step": frames,
                   "note": "oracle port of the reference's numpy/scipy path (oracle/ref_numpy.py), joblib threads over frames"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{frames} frames of {args.size}^2 per step, {args.steps} steps"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    from barc4dip_b200 import engine, parallel, synth
    from barc4dip_b200._lib import get_context
    from barc4dip_b200.pipeline import StackAnalyzer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # this image sets NCCL_DEBUG=VERSION and NCCL then writes its version banner to stdout; stdout carries the one JSON
        # line only (any other level the user asked for is kept)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]               # (WARN prints the banner as well)
        dist.init_process_group("nccl", device_id=dev)
    n, F = args.size, args.frames
    ctx = get_context(local)
    if args.batch:
        ctx.set_batch_frames(args.batch)

    # ---- synthetic stack, resident in HBM: SURVEY.md 8(d) C4 (sub-pixel random walk of one speckle + 1 % noise) ------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                  # sampled from before the clock ramp to the end of the timed region
    base = synth.speckle_frame(n, grain=6.0, seed=0)
    shifts = make_shifts(F, seed=2 + rank)
    stack = device_stack(base, shifts, 1234 + rank, dev)
    # reference frame = frame 0 of rank 0, broadcast over NCCL (the path's one exchange for tracking)
    ref = stack[0].clone()
    analyzer = StackAnalyzer((n, n), device=local, chunk_frames=F, want_maps=True, want_contrast=True)
    psd_out = torch.empty((F, n, n), dtype=torch.float32, device=dev)
    ac_out = torch.empty((F, n, n), dtype=torch.float32, device=dev)
    temporal = {}

    # One timed run = one stack. Start: the reference frame is broadcast (NCCL) and its conjugate spectrum built once;
    # rank 0's pilot mean of its first frames is broadcast as the common shift plane of the temporal power sums.
    # Every step analyses one batch of F frames per GPU and adds it to the temporal sums. End: the four float64
    # power-sum planes are all-reduced (NCCL) and finalised -- exactly how a 4000-frame stack is processed.
    def begin_stack():
        parallel.broadcast_reference(ref, src=0)
        analyzer.set_reference(ref)
        acc = engine.TemporalAccumulator(n, n, device=local)
        if rank == 0:
            acc.pilot(stack)
        else:
            acc.shift = torch.empty((n, n), dtype=torch.float32, device=dev)
        parallel.broadcast_reference(acc.shift, src=0)
        temporal["acc"] = acc

    def step():
        res = analyzer.run_device(stack, psd_out=psd_out, ac_out=ac_out, resolve_tails=False)
        temporal["acc"].update(stack)
        return res

    coll = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def end_stack(n_steps: int):
        acc = temporal["acc"]
        coll[0].record()
        parallel.allreduce_temporal(acc, world * F * n_steps)
        coll[1].record()
        return acc.finalize(return_device=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    begin_stack()
    # clock ramp: a fresh box idles at low SM clocks and needs ~1 s of load to reach its steady state; these passes are
    # neither warm-up steps nor timed steps (reported as config.prewarm_s)
    t_pre = time.perf_counter()
    n_pre = 0
    while time.perf_counter() - t_pre < args.prewarm:
        step()
        n_pre += 1
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        res = step()
    # a whole untimed stack before the timed one: the end of the stack (all-reduce, finalize) and a second start warm the
    # allocator's blocks and NCCL's channels the way any second stack of a session finds them
    coll[0].record()                    # (the ranks ran different numbers of untimed passes: no frame-count check here)
    parallel.allreduce_temporal(temporal["acc"])
    coll[1].record()
    temporal["acc"].finalize(return_device=True)
    begin_stack()
    step()
    end_stack(1)
    barrier()
    # sanity (untimed): the tracker against the generator's shifts on every frame, and against the oracle port of the
    # reference's phase_correlation on two sub-pixel frames
    tr = res["tracking"].cpu().numpy()
    whole = np.all(shifts == np.round(shifts), axis=1)
    err = float(np.max(np.abs(tr[whole, :2] - shifts[whole])))               # np.roll frames: known answers
    # sub-pixel frames: the reference's Taylor step is biased and adds the x term to dy: within a pixel
    err_sub = float(np.max(np.abs(tr[:, :2] - shifts)))
    if err > 0.05 or err_sub > 1.25:   # (|integer-peak error| <= 0.5 px plus the swapped sub-pixel terms, each <= ~0.5 px)
        raise SystemExit(f"tracking sanity check failed on rank {rank}: max |shift error| = {err:.3f} px (whole-pixel frames), "
                         f"{err_sub:.3f} px (all frames)")
    oracle_err = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import ref_numpy as orc
        host2 = stack[:4].cpu().numpy()
        full = (slice(0, n), slice(0, n))
        oracle_err = 0.0
        for t in (1, 3):
            want = orc.phase_correlation(host2[0], host2[t], slices_yx=full)
            oracle_err = max(oracle_err, float(np.max(np.abs(tr[t, :2] - np.asarray(want[:2])))))
        if oracle_err > 0.01:
            raise SystemExit(f"tracking differs from the oracle by {oracle_err:.4f} px (> 0.01 px)")
    snr_unresolved = int(torch.isnan(res["tracking"][:, 3]).sum().item())   # frames whose fused median would need the map-based path
    unresolved = int((res["n_valid"] < 0).sum().item())   # frames whose fused tail percentiles would need the exact fallback

    launches0 = ctx.launches
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_b, e_s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("b4d_timed")   # (ncu --nvtx --nvtx-include "b4d_timed/" profiles the timed region only)
    e0.record()
    begin_stack()                       # inside the timed region: both broadcasts + the reference spectrum
    e_b.record()
    for _ in range(args.steps):
        step()
    e_s.record()
    end_stack(args.steps)               # inside the timed region: all-reduce of the temporal sums + finalize
    e1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    prof = ctx.profile_end()
    launches = ctx.launches - launches0
    clock_info = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1), coll[0].elapsed_time(coll[1]), e0.elapsed_time(e_b), e_b.elapsed_time(e_s),
                       e_s.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms[0].item())
    coll_ms = float(ms[1].item())
    phases = {"begin_stack_ms": float(ms[2].item()), "steps_ms": float(ms[3].item()), "end_stack_ms": float(ms[4].item())}
    coll_bytes = 4 * n * n * 8
    collective = {"op": "all_reduce(sum, float64) of the 4 temporal power-sum planes + frame count", "bytes": coll_bytes,
                  "ms": coll_ms if world > 1 else 0.0,
                  "bus_gbs": (2.0 * (world - 1) / world) * coll_bytes / (coll_ms / 1e3) / 1e9 if world > 1 and coll_ms > 0 else None,
                  "broadcasts": "reference frame + temporal shift plane, float32, 2 x %.1f MB" % (n * n * 4 / MB),
                  "share_of_run": (coll_ms / total_ms) if world > 1 else 0.0}
    fps = world * F * args.steps / (total_ms / 1e3)

    # ---- end to end through the public API: pinned host stack -> results on the host ------------------------
    # e2e: what the reference's stack functions return (speckle_stack_stats / sharpness_stack_stats: per-frame scalar
    # tables) comes back to the host; the PSD and autocorrelation maps stay resident in HBM, as they do between the
    # reference's own stages. e2e_maps_to_host additionally drains both maps over PCIe.
    e2e = e2e_maps = e2e_u16 = None
    if not args.no_e2e:
        # (a step of the end-to-end legs = the first E frames of the stack: pinning the whole resident stack and its maps
        # would cost tens of GB of page-locked host memory for the same per-frame figure)
        E = max(1, min(F, args.e2e_frames))
        host = torch.empty((E, n, n), dtype=torch.float32, pin_memory=True)
        host.copy_(stack[:E])
        # detector-native frames: the same stack rounded to uint16 counts (half the PCIe bytes; widened on the device)
        host_u16 = torch.empty((E, n, n), dtype=torch.uint16, pin_memory=True)
        host_u16.copy_(stack[:E].clamp(0, 65535).round().to(torch.uint16))
        an2 = StackAnalyzer((n, n), device=local, chunk_frames=max(1, min(8, E // 4)), want_maps=True, want_contrast=True)
        ref_host = host[0].clone()

        def time_e2e(keep: bool, src=None):
            src = host if src is None else src

            def e2e_step():
                an2.set_reference(ref_host)
                return an2.run(src, keep_maps_on_device=keep, reuse_host_buffers=True)
            # two untimed calls: staging buffers, scratch arenas and the caching allocator's blocks for the result maps
            # are created on the first, reused from the second on (a result is released before the next call, as a
            # caller looping over stacks would)
            for _ in range(2):
                out = e2e_step()
                del out
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                out = e2e_step()
                del out
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return world * E * args.e2e_steps / float(dt.item())

        h2d_b, d2h_b = an2.bytes_per_frame()
        fps_tables = time_e2e(True)
        link_gbs = fps_tables / world * h2d_b / 1e9
        e2e = {"value": fps_tables, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b * E + n * n * 4),
               "d2h_bytes_per_step": int((d2h_b - 2 * n * n * 4) * E), "steps": args.e2e_steps, "frames_per_step_per_gpu": E,
               "h2d_gbs_per_gpu": link_gbs,
               "note": "StackAnalyzer.run on a pinned float32 host stack: every per-frame table (moments, sharpness, contrast, "
                       "grain, tracking) copied back, PSD + autocorrelation maps left in HBM. Bound by the host-to-device "
                       "link (h2d_gbs_per_gpu against ~55 GB/s of one PCIe 5 x16 link), not by the kernels"}
        fps_u16 = time_e2e(True, host_u16)
        e2e_u16 = {"value": fps_u16, "unit": UNIT, "h2d_bytes_per_step": int(n * n * 2 * E + n * n * 4),
                   "d2h_bytes_per_step": int((d2h_b - 2 * n * n * 4) * E), "steps": args.e2e_steps, "frames_per_step_per_gpu": E,
                   "h2d_gbs_per_gpu": fps_u16 / world * n * n * 2 / 1e9,
                   "note": "same call on the detector-native uint16 stack (b4d_cast_to_f32 widens on the device): half the "
                           "PCIe bytes per frame"}
        fps_maps = time_e2e(False)
        e2e_maps = {"value": fps_maps, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b * E + n * n * 4),
                    "d2h_bytes_per_step": int(d2h_b * E), "steps": args.e2e_steps, "frames_per_step_per_gpu": E,
                    "note": "same call, PSD + autocorrelation maps also copied to pinned host memory (PCIe-bound)"}

    # file -> results: the same uint16 frames as a gzip-4 chunked HDF5 stack (the reference's save_h5 layout), analysed by
    # io.stream.analyze_h5_stack -- stored chunks over PCIe, inflated by the GPU's decompression engine where there is one
    e2e_h5 = None
    if not args.no_e2e and world == 1:
        try:
            import tempfile
            from barc4dip_b200._lib import inflate_caps
            from barc4dip_b200.io import h5 as h5io
            from barc4dip_b200.io.stream import analyze_h5_stack
            Eh = min(E, 64)
            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, "stack.h5")
                h5io.save_h5(host_u16[:Eh].numpy(), path)
                an3 = StackAnalyzer((n, n), device=local, reference=ref_host, want_maps=False, want_contrast=True)
                mode = "device" if inflate_caps(local)[0] & 1 else "host"
                analyze_h5_stack(path, analyzer=an3, block_frames=32, inflate=mode)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    analyze_h5_stack(path, analyzer=an3, block_frames=32, inflate=mode)
                torch.cuda.synchronize()
                e2e_h5 = {"value": 2 * Eh / (time.perf_counter() - t0), "unit": UNIT, "inflate": mode, "frames": Eh,
                          "file_bytes_per_frame": os.path.getsize(path) / Eh, "raw_bytes_per_frame": n * n * 2,
                          "note": "HDF5 file (page cache) -> per-frame tables on the host; compressed chunks cross PCIe and are "
                                  "inflated on the device (b4d_inflate_batch + b4d_unchunk_to_f32) when inflate == 'device'"}
        except Exception as e:                                    # an extra figure must not cost the headline line
            e2e_h5 = {"error": repr(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the CUDA-event profile of the timed region --------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kern = {k: v for k, v in prof.items() if v[1] > 0}
    tot_kernel_ms = sum(v[0] for v in kern.values())
    dom = max(kern, key=lambda k: kern[k][0])
    dom_ms, dom_launches = kern[dom]
    frames_total = F * args.steps
    scale = (n * n) / (2048.0 * 2048.0)
    algo = ALGO_BYTES_PER_FRAME.get(dom, ALGO_BYTES_PER_FRAME["frame_reduce"]) * scale
    per_launch_frames = frames_total / dom_launches if dom in ("rows_fwd", "cols", "rows_inv", "frame_reduce") else None
    achieved = algo * frames_total / (dom_ms / 1e3) / 1e9
    # DRAM bytes the same kernel really moved (ncu --set full capture of this command, profiles/ncu_traffic.json)
    traffic = traffic_pf = traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        names = {"cols": "cols_kernel", "rows_inv": "rows_inv_kernel", "rows_fwd": "rows_fwd_kernel", "frame_reduce": "frame_reduce2_kernel"}
        ent = tj["kernels"].get(names.get(dom, dom))
        if ent and n == 2048:
            traffic_pf = float(ent["dram_bytes_per_frame"])
            traffic = traffic_pf * (per_launch_frames or F)
            traffic_src = tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_bytes_per_frame": traffic_pf, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "avg_launch_ms": dom_ms / dom_launches, "frames_per_launch": per_launch_frames,
                "algorithmic_bytes_per_frame": algo, "share_of_step": dom_ms / total_ms,
                "share_of_kernel_time": dom_ms / tot_kernel_ms,
                "share_note": "the tracker's and the autocorrelation branch run on two streams after the column pass: their "
                              "event spans overlap, so the kernel spans sum to more than the step; share_of_step = span / step time"}
    step_bytes = ALGO_BYTES_PER_FRAME["step"] * scale
    step_roof = {"bound": "hbm", "achieved": step_bytes * fps / world / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": step_bytes * fps / world / 1e9 / hbm_peak, "algorithmic_bytes_per_frame": step_bytes,
                 "note": "whole step per GPU: frame read once + PSD + autocorrelation maps written once"}
    lg = np.log2(n)
    fft_flops = 6144.0 * scale * (5.0 * n * lg) * (2048.0 / n)      # 6144 complex transforms of length n at 2048^2
    fp32 = {"achieved_tflops": fft_flops * fps / world / 1e12, "peak_tflops": 74.4,
            "frac": fft_flops * fps / world / 1e12 / 74.4, "note": "5 N log2 N flops, 1 forward + 2 inverse real 2-D FFTs per frame; "
            "peak = 148 SMs x 128 lanes x 2 x 1.965 GHz (non-tensor FP32)"}
    kernel_table = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                        "share_of_step": v[0] / total_ms} for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        cf = args.cpu_frames or max(8, min(32, cores))
        cfps, csec, cores = time_cpu(n, cf, 1, 0)
        cpu = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cf} frames of {n}^2, one pass of the oracle port (numpy/scipy, joblib threads) of the same per-frame work"}

    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fused stack pipeline (frame reductions + percentile contrast + psd2d map + autocorr2d map + grain + "
                               f"phase-correlation tracking vs broadcast reference + per-pixel temporal power sums, all-reduced at "
                               f"the end of the stack), {n}x{n} float32 frames (BASELINE configs[1..4] fused)",
                   "stack": "SURVEY 8(d) C4: one speckle Fourier-shifted along a 2-D random walk (sigma 0.3 px, +-20 px) + 1 % noise per frame",
                   "tracking_vs_oracle_max_px": oracle_err,
                   "frames_per_step_per_gpu": F, "frame": [n, n], "parallelism": f"frame-sharded x{world}",
                   "l2": f"inputs per step {F * n * n * 4 / MB:.0f} MB + {2 * F * n * n * 4 / MB:.0f} MB of maps written: larger than the 126 MB L2, no flush needed",
                   "internal_batch_frames": args.batch or "auto", "prewarm_s": args.prewarm,
                   "tail_percentile_frames_needing_fallback": unresolved, "tracker_median_frames_needing_fallback": snr_unresolved},
        "clocks": clock_info, "e2e": e2e, "e2e_uint16": e2e_u16, "e2e_maps_to_host": e2e_maps, "e2e_hdf5": e2e_h5, "gpu_launches": int(launches),
        "collective": collective, "phases": phases,
        "roofline": roofline, "step_roofline": step_roof, "fft_fp32": fp32, "kernels": kernel_table, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
