"""distribution_moments on the B200 path (drop-in for barc4dip.metrics.statistics, statistics.py:17-125)."""

from __future__ import annotations

import logging
import math

import numpy as np

from .. import engine
from .._lib import FR

logger = logging.getLogger(__name__)


def moments_from_row(row: np.ndarray, saturation_value: float | None) -> dict:
    """One row of the frame-reduction table -> the reference's 8-key moments dict."""
    n = float(row[FR["count"]])
    if n <= 0:
        raise ValueError("distribution_moments received no finite values.")
    mean = float(row[FR["mean"]])
    m2, m3, m4 = float(row[FR["m2"]]), float(row[FR["m3"]]), float(row[FR["m4"]])
    std = math.sqrt(m2)
    with np.errstate(divide="ignore", invalid="ignore"):
        skew = float(np.float64(m3) / np.float64(m2) ** 1.5)
        kurt = float(np.float64(m4) / np.float64(m2) ** 2 - 3.0)
    if std == 0.0:
        snr_db = float("inf") if mean > 0.0 else float("nan")
    else:
        q = mean / std
        snr_db = float(20.0 * math.log10(q)) if q > 0.0 else (float("-inf") if q == 0.0 else float("nan"))
    return {
        "mean": mean,
        "std": std,
        "variance": float(std * std),
        "skewness": skew,
        "kurtosis": kurt,
        "frac_zero": float(row[FR["nzero"]] / n),
        "frac_sat": float("nan") if saturation_value is None else float(row[FR["nsat"]] / n),
        "SNRdB": snr_db,
    }


def distribution_moments(image: np.ndarray, *, saturation_value: float | None = 65535.0, eps: float = 1e-6,
                         verbose: bool = False) -> dict:
    """Intensity distribution moments over the finite pixels (mean, std, variance, skewness, excess
    kurtosis, frac_zero, frac_sat, SNRdB).  Same signature, keys and error behaviour as the reference."""
    data = np.asarray(image)
    if data.ndim not in (1, 2):
        raise ValueError(f"Expected 1D or 2D array, got ndim={data.ndim}")
    if data.size == 0:
        raise ValueError("distribution_moments received an empty image.")
    frame = data.reshape(1, -1) if data.ndim == 1 else data
    table = engine.frame_reductions(engine.as_stack(frame), saturation_value=saturation_value, eps=eps)
    moments = moments_from_row(table[0], saturation_value)
    if verbose:
        logger.info("> moments: mean=%.0f | std=%.0f | var=%.0f | skew=%.2f | kurt=%.2f | SNR=%.2f dB | zero=%.6f | sat=%.6f",
                    moments["mean"], moments["std"], moments["variance"], moments["skewness"], moments["kurtosis"],
                    moments["SNRdB"], moments["frac_zero"], moments["frac_sat"])
    return moments
