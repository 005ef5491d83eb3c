"""
Deterministic synthetic inputs for tests, golden fixtures and the benchmark.

Test/bench infrastructure only: nothing here is on the product path. The
generators follow the recipe written down in SURVEY.md section 8(d) (speckle =
|IFFT(pupil * exp(2 pi i U))|^2 scaled to mean 1000, focus scan = Gaussian
blur, tracking stack = shifted copies of frame 0 + noise, temporal stack =
gamma speckle * flat + dark with dead pixels).

All functions take explicit integer seeds and return C-ordered float32.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "speckle_frame",
    "speckle_stack",
    "fourier_shift",
    "focus_scan_stack",
    "tracking_stack",
    "flatfield_case",
]


def speckle_frame(ny: int, nx: int | None = None, *, grain: float = 6.0,
                  seed: int = 0, mean: float = 1000.0) -> np.ndarray:
    """Fully developed speckle intensity of shape (ny, nx), float32.

    A circular pupil of radius 1/(2*grain) cycles/pixel filters a uniform random
    phase screen; the intensity is scaled to the requested mean.
    """
    nx = ny if nx is None else nx
    rng = np.random.default_rng(seed)
    phase = rng.uniform(0.0, 1.0, size=(ny, nx))
    fy = np.fft.fftfreq(ny)[:, None]
    fx = np.fft.fftfreq(nx)[None, :]
    pupil = (fx * fx + fy * fy) <= (0.5 / grain) ** 2
    field = np.fft.ifft2(pupil * np.exp(2j * np.pi * phase))
    inten = field.real ** 2 + field.imag ** 2
    inten *= mean / inten.mean()
    return np.ascontiguousarray(inten, dtype=np.float32)


def speckle_stack(t: int, ny: int, nx: int | None = None, *, grain: float = 6.0,
                  seed0: int = 0, mean: float = 1000.0) -> np.ndarray:
    """Stack of independent speckle frames, seeds seed0 .. seed0+t-1."""
    return np.stack([speckle_frame(ny, nx, grain=grain, seed=seed0 + k, mean=mean)
                     for k in range(t)], axis=0)


def fourier_shift(img: np.ndarray, dy: float, dx: float) -> np.ndarray:
    """Circularly translate img by (dy, dx) pixels (+dy down, +dx right), float32.

    Integer shifts are exactly np.roll; fractional ones use the Fourier shift
    theorem (band-limited interpolation).
    """
    a = np.asarray(img, dtype=np.float64)
    if float(dy).is_integer() and float(dx).is_integer():
        return np.roll(a, (int(dy), int(dx)), axis=(0, 1)).astype(np.float32)
    ny, nx = a.shape
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.fftfreq(nx)[None, :]
    ramp = np.exp(-2j * np.pi * (ky * dy + kx * dx))
    return np.fft.ifft2(np.fft.fft2(a) * ramp).real.astype(np.float32)


def focus_scan_stack(t: int, n: int, *, grain: float = 6.0, seed: int = 0,
                     noise_seed: int = 1) -> np.ndarray:
    """Config-2 style sharpness scan: one speckle blurred by sigma_k = 0.5 + |k - t/2| * 0.05 px."""
    base = speckle_frame(n, grain=grain, seed=seed).astype(np.float64)
    fb = np.fft.fft2(base)
    ky = np.fft.fftfreq(n)[:, None]
    kx = np.fft.fftfreq(n)[None, :]
    k2 = kx * kx + ky * ky
    rng = np.random.default_rng(noise_seed)
    out = np.empty((t, n, n), dtype=np.float32)
    for k in range(t):
        sigma = 0.5 + abs(k - t // 2) * 0.05
        blurred = np.fft.ifft2(fb * np.exp(-2.0 * (np.pi * sigma) ** 2 * k2)).real
        blurred = np.clip(blurred, 0.0, None)
        noisy = blurred + rng.standard_normal((n, n)) * np.sqrt(blurred + 1.0)
        out[k] = noisy
    return out


def tracking_stack(t: int, n: int, *, grain: float = 6.0, seed: int = 0,
                   walk_seed: int = 2, noise_seed: int = 3, step_sigma: float = 0.3,
                   clip: float = 20.0, noise_frac: float = 0.01,
                   integer_every: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """Config-4 style stack: frame t = a clean speckle translated along a 2-D random walk + 1 % noise
    (independent noise on every frame, frame 0 included: a noise-free band-limited reference would leave the
    phase of its empty spectral bins to FFT rounding noise and make the tracker's output ill-conditioned).

    Returns (stack (t, n, n) float32, shifts (t, 2) float64 as (dy, dx)); shifts[0] = (0, 0).
    If integer_every > 0 every such frame gets an integer (np.roll) shift: a known-answer case.
    """
    base = speckle_frame(n, grain=grain, seed=seed)
    wrng = np.random.default_rng(walk_seed)
    nrng = np.random.default_rng(noise_seed)
    steps = wrng.normal(0.0, step_sigma, size=(t, 2))
    steps[0] = 0.0
    shifts = np.clip(np.cumsum(steps, axis=0), -clip, clip)
    if integer_every > 0:
        shifts[::integer_every] = np.round(shifts[::integer_every] * 4.0)
        shifts[0] = 0.0
    sigma = noise_frac * float(base.mean())
    out = np.empty((t, n, n), dtype=np.float32)
    out[0] = base + nrng.normal(0.0, sigma, size=(n, n)).astype(np.float32)   # the reference frame is noisy too
    for k in range(1, t):
        fr = fourier_shift(base, shifts[k, 0], shifts[k, 1])
        out[k] = fr + nrng.normal(0.0, sigma, size=(n, n)).astype(np.float32)
    return out, shifts


def flatfield_case(t: int, n: int, *, seed: int = 5, dead_frac: float = 1e-4
                   ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Config-5 style raw stack with its flat and dark: (images (t,n,n), flat (n,n), dark (n,n)).

    raw = gamma-speckle(mean 1000) * (1 + 0.1 * smooth(x, y)) + dark, dark = 100 + N(0, 2);
    a fraction dead_frac of pixels has flat == dark (denominator <= eps -> bad-pixel mask).
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n), indexing="ij")
    gain = 1.0 + 0.1 * np.cos(1.3 * xx) * np.sin(0.7 * yy + 0.2)
    dark = 100.0 + rng.normal(0.0, 2.0, size=(n, n))
    flat = 5000.0 * gain + dark
    n_dead = max(1, int(round(dead_frac * n * n)))
    dead = rng.choice(n * n, size=n_dead, replace=False)
    flat.ravel()[dead] = dark.ravel()[dead]
    sig = rng.exponential(1000.0, size=(t, n, n))
    raw = sig * gain[None] + dark[None]
    return (raw.astype(np.float32), flat.astype(np.float32), dark.astype(np.float32))
