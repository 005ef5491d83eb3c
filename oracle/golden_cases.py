"""
Seeded inputs shared by ``make_golden.py`` (which runs the real reference on them) and the
tests (which regenerate them and compare the oracle / the CUDA path with the stored
reference outputs).  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import numpy as np

from barc4dip_b200 import synth


def frame_cases() -> dict[str, np.ndarray]:
    """2-D frames for the per-frame functions (F1-F4, F7-F13)."""
    sq = synth.speckle_frame(256, grain=4.0, seed=11)
    rect = synth.speckle_frame(128, 256, grain=5.0, seed=12)
    odd = synth.speckle_frame(150, 200, grain=4.0, seed=13)          # non power of two (oracle only)
    u16 = np.clip(synth.speckle_frame(128, grain=3.0, seed=14) * 8.0, 0, 65535).astype(np.uint16)
    u16.ravel()[::997] = 65535                                        # saturated pixels
    u16.ravel()[5::1013] = 0                                          # zero pixels
    smooth = synth.focus_scan_stack(3, 256, grain=6.0, seed=15, noise_seed=16)[0]
    big = synth.speckle_frame(512, grain=6.0, seed=17)
    return {"sq256": sq, "rect128x256": rect, "odd150x200": odd, "u16_128": u16,
            "blur256": smooth, "sq512": big}


def nan_frame() -> np.ndarray:
    a = synth.speckle_frame(128, grain=4.0, seed=18).copy()
    a[3, 7] = np.nan
    a[64, 64] = np.inf
    a[127, 0] = -np.inf
    return a


def tracking_cases() -> dict[str, dict]:
    """(template, image, slices) triples for phase correlation (F5/F6)."""
    n = 256
    base = synth.speckle_frame(n, grain=4.0, seed=21)
    rng = np.random.default_rng(22)
    full = (slice(0, n), slice(0, n))
    cases: dict[str, dict] = {}

    def add(name, dy, dx, noise=0.0, tpl=None, slices=full, subpixel=True):
        # Cases with noise carry independent noise on BOTH frames (as any detector frame does). A noise-free
        # band-limited frame has empty spectral bins whose whitened phase is pure FFT rounding noise, which
        # makes the result implementation-defined at the 0.03 px level (float32 vs float64 numpy disagree).
        img = synth.fourier_shift(base, dy, dx)
        ref = base
        if noise:
            sigma = noise * float(base.mean())
            img = img + rng.normal(0.0, sigma, size=img.shape).astype(np.float32)
            ref = base + rng.normal(0.0, sigma, size=base.shape).astype(np.float32)
        t = ref[slices] if tpl is None else tpl
        cases[name] = {"template": np.ascontiguousarray(t), "image": np.ascontiguousarray(img),
                       "slices": slices, "subpixel": subpixel, "true_shift": (dy, dx), "noisy": bool(noise)}

    add("roll_2_m3", 2, -3)
    add("sub_p3_m1", 0.3, -0.1)
    add("sub_m5p25_7p6_noise", -5.25, 7.6, noise=0.01)
    add("sub_12p5_m9p75_noise5", 12.5, -9.75, noise=0.05)
    add("nosubpixel", 1.4, 2.6, subpixel=False)
    add("sub_p3_m1_noise", 0.3, -0.1, noise=0.01)
    add("nosubpixel_noise", 1.4, 2.6, noise=0.02, subpixel=False)
    add("zero_shift", 0, 0, noise=0.01)
    # odd ROI, centred, slices_yx=None
    c = n // 2
    roi = (slice(c - 75, c + 76), slice(c - 75, c + 76))
    add("roi151_centered", 3.3, -2.2, noise=0.01, slices=roi)
    cases["roi151_centered"]["slices"] = None
    # off-centre ROI with explicit slices
    roi2 = (slice(20, 20 + 181), slice(40, 40 + 161))
    add("roi181x161_offcentre", -1.7, 4.4, noise=0.01, slices=roi2)
    # integer dtype inputs (converted to float32 by the tracker)
    img_i = np.clip(synth.fourier_shift(base, 4, 6), 0, 65535).astype(np.uint16)
    cases["uint16_roll_4_6"] = {"template": np.clip(base, 0, 65535).astype(np.uint16), "image": img_i,
                                "slices": full, "subpixel": True, "true_shift": (4, 6), "noisy": False}
    # rectangular frame
    rb = synth.speckle_frame(128, 256, grain=4.0, seed=23)
    cases["rect_roll_m6_9"] = {"template": rb, "image": synth.fourier_shift(rb, -6, 9),
                               "slices": (slice(0, 128), slice(0, 256)), "subpixel": True,
                               "true_shift": (-6, 9), "noisy": False}
    rbn = rb + rng.normal(0.0, 20.0, size=rb.shape).astype(np.float32)
    cases["rect_sub_noise"] = {"template": rbn, "image": synth.fourier_shift(rb, 3.4, -8.3)
                               + rng.normal(0.0, 20.0, size=rb.shape).astype(np.float32),
                               "slices": (slice(0, 128), slice(0, 256)), "subpixel": True,
                               "true_shift": (3.4, -8.3), "noisy": True}
    return cases


def flatfield_inputs():
    return synth.flatfield_case(4, 128, seed=31, dead_frac=2e-3)


def flatfield_repair_inputs():
    """3 frames of 96 x 80 with 1 % dead pixels, some of them adjacent to one another and on the borders / corners
    (the median filter's 'reflect' border and the "bad neighbours count as 0" rule both get exercised)."""
    raw, flat, dark = synth.flatfield_case(3, 96, seed=33, dead_frac=1e-2)
    raw, flat, dark = raw[:, :, :80].copy(), flat[:, :80].copy(), dark[:, :80].copy()
    for (y, x) in ((0, 0), (0, 1), (1, 0), (95, 79), (95, 40), (50, 0), (50, 1), (51, 1), (20, 79), (60, 30), (60, 31), (61, 30)):
        flat[y, x] = dark[y, x]                       # den = 0 -> bad
    return raw, flat, dark


def temporal_inputs():
    raw, flat, dark = synth.flatfield_case(24, 64, seed=41, dead_frac=2e-3)
    return raw, flat, dark


def tile_cases() -> dict[str, np.ndarray]:
    """Frames for the 3x3 / 9x9 tiling executor (metrics/common.py:278-378): 512^2 -> tiles_3x3 with 170 / 171 px tiles;
    1170 x 1200 -> subtiles_9x9 with 130 px x 133 / 134 px sub-tiles (none a power of two)."""
    return {"t3_512": synth.speckle_frame(512, grain=6.0, seed=17),
            "t9_1170x1200": synth.speckle_frame(1170, 1200, grain=5.0, seed=19)}


SHARPNESS_TILE_GROUPS = ("stats", "gradient", "laplacian", "spectral", "autocorrelation")


def template_cases() -> dict[str, dict]:
    """(template, image, slices) triples for template_matching (signal/tracking.py:82-188, opencv backend)."""
    n = 256
    base = synth.speckle_frame(n, grain=4.0, seed=61)
    rng = np.random.default_rng(62)
    noise = lambda: (0.02 * float(base.mean()) * rng.standard_normal((n, n))).astype(np.float32)
    cases = {}
    sl = (slice(100, 125), slice(140, 165))                      # 25 x 25 ROI
    cases["roll_25"] = dict(template=base[sl].copy(), image=np.roll(base, (3, -5), axis=(0, 1)) + noise(), slices=sl, subpixel=True)
    sub = synth.fourier_shift(base, 1.37, -2.62) if hasattr(synth, "fourier_shift") else None
    if sub is None:
        F = np.fft.fft2(base.astype(np.float64))
        ky = np.fft.fftfreq(n)[:, None]; kx = np.fft.fftfreq(n)[None, :]
        sub = np.real(np.fft.ifft2(F * np.exp(-2j * np.pi * (ky * 1.37 + kx * -2.62)))).astype(np.float32)
    cases["subpx_25"] = dict(template=base[sl].copy(), image=sub + noise(), slices=sl, subpixel=True)
    cases["subpx_25_nosub"] = dict(template=base[sl].copy(), image=sub + noise(), slices=sl, subpixel=False)
    sl2 = (slice(60, 161), slice(30, 105))                       # 101 x 75 ROI
    cases["rect_101x75"] = dict(template=base[sl2].copy(), image=np.roll(base, (-7, 4), axis=(0, 1)) + noise(), slices=sl2, subpixel=True)
    cases["centered_31"] = dict(template=base[112:143, 112:143].copy(), image=np.roll(base, (2, 2), axis=(0, 1)) + noise(),
                                slices=None, subpixel=True)
    u16 = np.clip(base * 20.0, 0, 65535).astype(np.uint16)
    cases["u16_25"] = dict(template=u16[sl].copy(), image=np.roll(u16, (1, 6), axis=(0, 1)), slices=sl, subpixel=True)
    return cases


def stack_tracking_case() -> np.ndarray:
    """4 frames of 256^2 drifting by a random walk (speckle_stack_stats, template tracker on the 3x3 ROI grid)."""
    stack, _ = synth.tracking_stack(4, 256, grain=8.0, seed=51, step_sigma=0.8)
    return stack


def signal1d_cases() -> dict[str, np.ndarray]:
    """1-D signals for fft1d / psd1d / xcorr1d / autocorr1d: a power-of-two length and one that is not."""
    rng = np.random.default_rng(71)
    t = np.arange(1024)
    a = (np.sin(2 * np.pi * t / 37.0) + 0.3 * rng.standard_normal(1024) + 2.0).astype(np.float32)
    return {"n1024": a, "n300": a[100:400].copy()}
