"""
read_tiff with the reference's contract (io/tiff.py:19-70): one file -> the image as PIL decodes it, a sequence of
files -> one (N, H, W) stack. PIL is the decoder on both sides; it is imported on first use so that the package loads
without it. save_tiff is not built (the reference rescales through utils/dtype.py:to_uint16, a display conversion that is
not part of the stack path).
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np


def read_tiff(image_path: str | Sequence[str]) -> np.ndarray:
    from PIL import Image

    if isinstance(image_path, str):
        with Image.open(image_path) as img:
            return np.array(img)
    if not isinstance(image_path, Sequence):
        raise TypeError("image_path must be a str or a sequence of str")
    if len(image_path) == 0:
        raise ValueError("image_path sequence is empty")
    frames = []
    for path in image_path:
        if not isinstance(path, str):
            raise TypeError("All elements of image_path must be strings")
        with Image.open(path) as img:
            arr = np.array(img)
        if frames and arr.shape != frames[0].shape:
            raise ValueError(f"Inconsistent image shapes in stack: expected {frames[0].shape}, got {arr.shape} for '{path}'")
        frames.append(arr)
    return np.stack(frames, axis=0)
