"""
Stack-level metric blocks: every function takes an HBM-resident (T, ny, nx) float32 CUDA tensor and
returns a dict of (T,) numpy arrays with the reference's key names.  The per-frame drop-in
functions in ``metrics/`` call these with T = 1; the stack aggregators call them once per batch.

Reference definitions: metrics/statistics.py:60-110, metrics/sharpness.py:440-465, :510-525, :581-629,
:681-744, metrics/speckles.py:530-585, :636-654, :740-805.
"""

from __future__ import annotations

import math

import numpy as np

from . import engine
from ._lib import FR, SP, require_cuda

INV_E = 1.0 / math.e


def moments_block(table: np.ndarray, saturation_value) -> dict:
    """Frame-reduction table (T, FR_NCOLS) -> dict of (T,) arrays with distribution_moments' keys."""
    n = table[:, FR["count"]]
    if np.any(n <= 0):
        raise ValueError("distribution_moments received no finite values.")
    mean, m2, m3, m4 = (table[:, FR[k]] for k in ("mean", "m2", "m3", "m4"))
    std = np.sqrt(m2)
    with np.errstate(divide="ignore", invalid="ignore"):
        skew = m3 / m2 ** 1.5
        kurt = m4 / m2 ** 2 - 3.0
        q = mean / std
        snr = np.where(q > 0, 20.0 * np.log10(np.where(q > 0, q, 1.0)), np.where(q == 0, -np.inf, np.nan))
    snr = np.where(std == 0.0, np.where(mean > 0.0, np.inf, np.nan), snr)
    return {
        "mean": mean.copy(), "std": std, "variance": std * std, "skewness": skew, "kurtosis": kurt,
        "frac_zero": table[:, FR["nzero"]] / n,
        "frac_sat": np.full(n.shape, np.nan) if saturation_value is None else table[:, FR["nsat"]] / n,
        "SNRdB": snr,
    }


def gradient_block(table: np.ndarray, eps: float = 1e-12) -> dict:
    n = table[:, FR["count"]]
    if np.any(n <= 0):
        raise ValueError("tenengrad received image with no finite values.")
    ex, ey = table[:, FR["sgx2"]] / n, table[:, FR["sgy2"]] / n
    return {"tenengrad": ex + ey, "ex": ex, "ey": ey, "re": ex / (ey + float(eps))}


def laplacian_block(table: np.ndarray) -> dict:
    n = table[:, FR["count"]]
    if np.any(n <= 0):
        raise ValueError("laplacian_variance received image with no finite values.")
    m = table[:, FR["slap"]] / n
    return {"laplacian_variance": table[:, FR["slap2"]] / n - m * m}


def amplitude_block(stack, table: np.ndarray | None = None) -> dict:
    """visibility = nanstd / nanmean; contrast from the exact 0.05 / 99.95 percentiles (linear interpolation)."""
    if table is None:
        table = engine.frame_reductions(stack, saturation_value=None)
    n = table[:, FR["count"]]
    n_inf = table[:, FR["npix"]] - n - table[:, FR["nnan"]]
    mu = np.where(n_inf > 0, np.nan, table[:, FR["mean"]])
    if np.any(~np.isfinite(mu)) or np.any(mu <= 0.0) or np.any(n <= 0):
        raise ValueError("Mean intensity must be positive and finite.")
    vis = np.sqrt(table[:, FR["m2"]]) / mu
    vals, nv = engine.select_quantiles(stack, [0.05 / 100.0, 99.95 / 100.0])
    T = vals.shape[0]
    contrast = np.empty(T)
    for t in range(T):
        vmin = engine.quantile_from_bracket(vals[t, 0], vals[t, 1], int(nv[t]), 0.05 / 100.0)
        vmax = engine.quantile_from_bracket(vals[t, 2], vals[t, 3], int(nv[t]), 99.95 / 100.0)
        den = vmax + vmin
        if not np.isfinite(den) or den <= 0.0:
            raise ValueError("Invalid percentile range for Michelson contrast.")
        contrast[t] = (vmax - vmin) / den
    return {"visibility": vis, "contrast": contrast}


def pad_to_square_stack(stack, means: np.ndarray):
    """Centre every frame in an (N, N) frame filled with that frame's mean (geometry/masks.py:11-56). Device copy."""
    torch = require_cuda()
    T, H, W = stack.shape
    if H == W:
        return stack
    N = max(H, W)
    out = torch.empty((T, N, N), dtype=torch.float32, device=stack.device)
    out[:] = torch.as_tensor(means, dtype=torch.float32, device=stack.device).view(T, 1, 1)
    y0, x0 = (N - H) // 2, (N - W) // 2
    out[:, y0:y0 + H, x0:x0 + W] = stack
    return out


def _frame_means(stack, table=None) -> np.ndarray:
    if table is None:
        table = engine.frame_reductions(stack, saturation_value=None)
    n_bad = table[:, FR["npix"]] - table[:, FR["count"]]
    return np.where(n_bad > 0, np.nan, table[:, FR["mean"]])


def grain_block(stack, *, fraction: float = INV_E, standardize: bool = False, table=None, return_map: bool = False):
    """lx, ly, leq, r per frame from the peak-normalised autocorrelation (+ the device map if asked)."""
    sq = pad_to_square_stack(stack, _frame_means(stack, table))
    ac, g = engine.autocorr2d(sq, remove_mean=True, standardize=standardize, normalize_peak=True,
                              want_map=return_map, want_grain=True, fraction=fraction)
    out = {"lx": g[:, 0], "ly": g[:, 1], "leq": g[:, 2], "r": g[:, 3]}
    return (out, ac) if return_map else out


def bandwidth_block(stack, *, table=None) -> dict:
    sq = pad_to_square_stack(stack, _frame_means(stack, table))
    n = sq.shape[-1]
    _, sp = engine.psd2d(sq, scale_factor=1.0 / (float(n) * float(n)), sub_mean=True, zero_dc=True,
                         want_map=False, want_spectral=True)
    return bandwidth_from_sums(sp)


def bandwidth_from_sums(sp: np.ndarray) -> dict:
    """bandwidth() metrics (metrics/speckles.py:771-805) from a spectral table (T, SP_NCOLS); invariant to the PSD's scale."""
    total = sp[:, SP["total"]]
    if np.any(~np.isfinite(total)) or np.any(total <= 0.0):
        raise ValueError("PSD energy is not positive/finite after mean/DC removal.")
    sfx, sfy = np.sqrt(sp[:, SP["fx2"]] / total), np.sqrt(sp[:, SP["fy2"]] / total)
    with np.errstate(divide="ignore"):
        rf = np.where(sfy != 0.0, sfx / sfy, np.inf)
    return {"feq": np.sqrt((sp[:, SP["fx2"]] + sp[:, SP["fy2"]]) / total), "f95": sp[:, SP["f95"]],
            "sig_fx": sfx, "sig_fy": sfy, "rf": rf, "spr": total * total / sp[:, SP["p2"]]}


def spectral_entropy_block(stack) -> dict:
    T, ny, nx = stack.shape
    _, sp = engine.psd2d(stack, scale_factor=1.0, sub_mean=True, zero_dc=True, want_map=False, want_spectral=True)
    return entropy_from_sums(sp, ny, nx)


def entropy_from_sums(sp: np.ndarray, ny: int, nx: int) -> dict:
    """spectral_entropy() (metrics/sharpness.py:611-621) from a spectral table; invariant to the PSD's scale."""
    s = sp[:, SP["all"]]
    if np.any(~np.isfinite(s)) or np.any(s <= 0.0):
        raise ValueError("PSD sum is non-positive; cannot compute spectral entropy.")
    m = ny * nx - 1
    if m < 2:
        raise ValueError("Insufficient number of spectral bins to compute normalized entropy.")
    h = np.log(s) - sp[:, SP["plogp"]] / s
    return {"spectral_entropy": h / math.log(float(m))}


def inverse_autocorr_block(stack, *, fraction: float = INV_E, table=None) -> dict:
    return inverse_from_grain(grain_block(stack, fraction=fraction, standardize=True, table=table))


def inverse_from_grain(g: dict) -> dict:
    with np.errstate(divide="ignore"):
        inv = lambda v: np.where(v != 0.0, 1.0 / v, np.inf)
        return {"sx": inv(g["lx"]), "sy": inv(g["ly"]), "seq": inv(g["leq"]),
                "r": np.where(g["ly"] != 0.0, g["lx"] / g["ly"], np.inf)}


def fused_available(stack) -> bool:
    """Square frames with a power-of-two side in [128, 2048]: the fused pass serves every FFT-based block at once."""
    ny, nx = int(stack.shape[-2]), int(stack.shape[-1])
    return ny == nx and 128 <= ny <= 2048 and (ny & (ny - 1)) == 0


class FusedBlocks:
    """One fused pass over an HBM-resident stack (b4d_stack_pipeline_ref without the tracker): frame reductions, tail
    order statistics, autocorrelation with grain widths and the spectral sums of bandwidth() / spectral_entropy() all
    come from ONE streaming read and ONE forward 2-D FFT per frame, where composing the per-metric blocks above runs a
    forward transform per FFT-based metric (the reference's aggregators call the metric functions one after the other,
    metrics/speckles.py:168-190, metrics/sharpness.py:183-211). The per-group dicts have the blocks' formats."""

    Q = (0.05 / 100.0, 99.95 / 100.0)

    def __init__(self, stack, *, saturation_value, eps: float, keep_map: bool = False, want_tails: bool = True,
                 want_spectral: bool = True):
        res = engine.stack_pipeline(stack, saturation_value=saturation_value, eps=eps, want_psd=False, want_autocorr=keep_map,
                                    want_grain=True, want_tracking=False, want_spectral=want_spectral,
                                    tail_quantiles=self.Q if want_tails else None)
        self.stack, self.sat = stack, saturation_value
        self.table = res["reductions"].cpu().numpy()
        self.ac = res["autocorr"]
        g = res["grain"].cpu().numpy()
        self._grain = {"lx": g[:, 0], "ly": g[:, 1], "leq": g[:, 2], "r": g[:, 3]}
        self._sp = res["spectral"].cpu().numpy() if want_spectral else None
        self._quant, self._nv = res["quantiles"], res["n_valid"]

    def moments(self) -> dict:
        return moments_block(self.table, self.sat)

    def amplitude(self) -> dict:
        table = self.table
        n = table[:, FR["count"]]
        n_inf = table[:, FR["npix"]] - n - table[:, FR["nnan"]]
        mu = np.where(n_inf > 0, np.nan, table[:, FR["mean"]])
        if np.any(~np.isfinite(mu)) or np.any(mu <= 0.0) or np.any(n <= 0):
            raise ValueError("Mean intensity must be positive and finite.")
        vis = np.sqrt(table[:, FR["m2"]]) / mu
        # frames the fused tail collection did not resolve (n_valid == -1) are redone by the exact stand-alone select
        engine.resolve_tail_quantiles(self.stack, self._quant, self._nv, *self.Q)
        q, nv = self._quant.cpu().numpy(), self._nv.cpu().numpy()
        contrast = np.empty(q.shape[0])
        for t in range(q.shape[0]):
            vmin = engine.quantile_from_bracket(q[t, 0], q[t, 1], int(nv[t]), self.Q[0])
            vmax = engine.quantile_from_bracket(q[t, 2], q[t, 3], int(nv[t]), self.Q[1])
            den = vmax + vmin
            if not np.isfinite(den) or den <= 0.0:
                raise ValueError("Invalid percentile range for Michelson contrast.")
            contrast[t] = (vmax - vmin) / den
        return {"visibility": vis, "contrast": contrast}

    def grain(self) -> dict:
        return dict(self._grain)

    def bandwidth(self) -> dict:
        return bandwidth_from_sums(self._sp)

    def entropy(self) -> dict:
        return entropy_from_sums(self._sp, int(self.stack.shape[-2]), int(self.stack.shape[-1]))

    def inverse_autocorr(self) -> dict:
        return inverse_from_grain(self._grain)


def eigenvalues_block(stack, *, k: int = 5, eps: float = 1e-30) -> dict:
    """STA2 focus measure (metrics/sharpness.py:752-861) for every frame of a (T, H, W) device stack.

    Outside the hot path: a dense symmetric eigenproblem, served by library code -- a float64 Gram
    matrix J J^T (or J^T J, whichever is smaller; cuBLAS) and cuSOLVER's symmetric eigensolver through torch.linalg --
    on the device, so that metrics="all" needs no host round trip of the stack. The eigenvalues of the Gram matrix are
    the squared singular values the reference takes from numpy.linalg.svd."""
    torch = require_cuda()
    T, H, W = stack.shape
    k_eff = int(k)
    if k_eff < 1:
        raise ValueError("k must be >= 1.")
    denom = float(H * W - 1)
    if denom <= 0.0:
        raise ValueError("eigenvalues requires at least 2 pixels (M*N >= 2).")
    n = min(H, W)
    chunk = max(1, (1 << 25) // (H * W))                  # frames per batched eigensolve: ~0.8 GB of float64 working set
    top = torch.zeros((T, max(2, min(k_eff, n))), dtype=torch.float64, device=stack.device)
    for t0 in range(0, T, chunk):
        x = stack[t0:t0 + chunk].to(torch.float64)
        if not bool(torch.isfinite(x).all()):
            raise ValueError("eigenvalues requires all values to be finite.")
        energy = torch.sqrt((x * x).sum(dim=(1, 2), keepdim=True))
        if bool((energy <= 0.0).any()) or not bool(torch.isfinite(energy).all()):
            raise ValueError("eigenvalues cannot normalize an all-zero image.")
        J = x / energy
        J = J - J.mean(dim=(1, 2), keepdim=True)
        G = J @ J.transpose(1, 2) if H <= W else J.transpose(1, 2) @ J
        ev = torch.linalg.eigvalsh(G).flip(-1).clamp_min(0.0) / denom          # descending
        m = min(top.shape[1], ev.shape[1])
        top[t0:t0 + chunk, :m] = ev[:, :m]
    top = top.cpu().numpy()
    k_use = min(k_eff, n)
    e1 = top[:, 0]
    e2 = top[:, 1] if n >= 2 else np.zeros(T)
    return {"eigenvalues": top[:, :k_use].sum(axis=1), "e1": e1, "e2": e2, "re": e1 / (e2 + float(eps))}
