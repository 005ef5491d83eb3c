"""
read_image / write_image: the reference's extension dispatchers (io/rw.py:64-148, :151-189): every format it reads
(HDF5, TIFF, EDF) and HDF5 output.

TIFF output is recognised (same extension tables, same argument checks in the same order) and then refused with a
ValueError naming the format: the display rescaling behind save_tiff (utils/dtype.py:to_uint16) is not part of the stack
path. EDF output is refused as in the reference.
"""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import numpy as np

from .edf import read_edf
from .h5 import read_h5, save_h5
from .tiff import read_tiff

_READ_EXTS = {"tif": "tiff", "tiff": "tiff", "edf": "edf", "h5": "h5", "hdf5": "h5"}
_WRITE_EXTS = {"tif": "tiff", "tiff": "tiff", "h5": "h5", "hdf5": "h5", "edf": "edf"}


def _ext_of(path: str) -> str:
    suffix = Path(path).suffix
    if suffix == "":
        raise ValueError("Cannot infer file extension from path (no suffix). Provide file_extension explicitly.")
    return suffix.lower().lstrip(".")


def read_image(image_path: str | Sequence[str], *, file_extension: str | None = None, image_number: int | None = None,
               mean: bool = False, verbose: bool = False) -> np.ndarray:
    forced = file_extension.lower().lstrip(".") if file_extension else None
    if isinstance(image_path, str):
        ext = forced or _ext_of(image_path)
    elif isinstance(image_path, Sequence):
        if len(image_path) == 0:
            raise ValueError("image_path sequence is empty")
        if forced:
            ext = forced
        else:
            exts = [_ext_of(p) for p in image_path]
            if any(e != exts[0] for e in exts):
                raise ValueError(f"Mixed file extensions in image_path sequence: {sorted(set(exts))}")
            ext = exts[0]
    else:
        raise TypeError("image_path must be a str or a sequence of str")
    if not isinstance(image_path, str) and image_number is not None:
        raise ValueError("image_number is only supported when image_path is a single file (str)")
    kind = _READ_EXTS.get(ext)
    if kind is None:
        raise ValueError(f"Unsupported read extension: '{ext}'")
    if kind != "h5":
        if image_number is not None:
            raise ValueError("image_number is only supported for HDF5 stacks (single-file .h5/.hdf5).")
        data = read_edf(image_path) if kind == "edf" else read_tiff(image_path)
    else:
        data = read_h5(image_path, image_number=image_number)
    if mean and data.ndim == 3:
        data = data.mean(axis=0)
        if verbose:
            print("Collapsed 3D stack to mean image along axis 0.")
    if verbose:
        n_img, (h, w) = (1, data.shape) if data.ndim == 2 else (data.shape[0], data.shape[1:])
        print(f"> {n_img} image(s) ({h} x {w}), {data.nbytes / 1024 ** 3:.2f} Gb in memory")
    return data


def write_image(data: np.ndarray, output_path: str | Path, *, file_extension: str | None = None,
                verbose: bool = False) -> None:
    if not isinstance(data, np.ndarray):
        raise TypeError("data must be a numpy.ndarray")
    out = Path(output_path)
    ext = file_extension.lower().lstrip(".") if file_extension else _ext_of(str(out))
    kind = _WRITE_EXTS.get(ext)
    if kind is None:
        raise ValueError(f"Unsupported write extension: '{ext}'")
    if kind == "edf":
        raise ValueError("Writing EDF is not supported (legacy read-only format).")
    if kind != "h5":
        raise ValueError("TIFF output is not built in barc4dip_b200 (HDF5 only)")
    save_h5(data, out)
    if verbose:
        print(f"> image saved to {out}")
