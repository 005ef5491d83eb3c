"""
read_edf with the reference's contract (io/edf.py:18-91): frame `index` of an ESRF data format file, cast to `dtype`
(float32 by default); a sequence of files -> one (N, H, W) stack of 2-D frames.

The reference delegates to a 1300-line legacy parser (io/uti_EdfFile.py) that also wraps TIFF, MarCCD, Pilatus CBF, SPE
and ADSC files; this is a reader of plain EDF only, written from the format itself: a file is a sequence of images,
each an ASCII header `{ key = value ; ... }` closed by a line holding `}` and followed at once by `Size` bytes of
binary data; `Dim_1` is the fastest axis, `DataType` / `ByteOrder` name the element type (io/uti_EdfFile.py:320-404,
:1106-1123 for the type names it accepts). Pinned against arrays the reference itself read from the same bytes
(tests/golden/edf.npz, oracle/make_golden_io.py).
"""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import numpy as np

_TYPES = {"SIGNEDBYTE": "i1", "UNSIGNEDBYTE": "u1", "SIGNEDSHORT": "i2", "UNSIGNEDSHORT": "u2", "SIGNEDINTEGER": "i4",
          "UNSIGNEDINTEGER": "u4", "SIGNEDLONG": "i4", "UNSIGNEDLONG": "u4", "SIGNED64": "i8", "UNSIGNED64": "u8",
          "FLOATVALUE": "f4", "FLOAT": "f4", "DOUBLEVALUE": "f8"}
_STATIC = ("SIZE", "DIM_1", "DIM_2", "DIM_3", "DATATYPE", "BYTEORDER")


def _images(raw: bytes) -> list[dict]:
    """Header fields and data position of every image in the file, in order."""
    out, pos, n = [], 0, len(raw)
    cur = None
    while pos < n:
        end = raw.find(b"\n", pos)
        end = n if end < 0 else end + 1
        line = raw[pos:end].decode("latin-1")
        pos = end
        if "{\n" in line or "{\r\n" in line:
            cur = {}
        if cur is not None and "=" in line:
            key, rest = line.split("=", 1)
            if key.strip().upper() in _STATIC:
                cur[key.strip().upper()] = rest.split(";", 1)[0].strip()
        if cur is not None and ("}\n" in line or "}\r" in line):
            if "SIZE" not in cur:
                raise TypeError("EdfFile: Image doesn't have size information")
            size = int(cur["SIZE"])
            if size <= 0:                                            # an empty trailing image ends the list
                break
            for need, what in (("DIM_1", "dimension"), ("DATATYPE", "datatype"), ("BYTEORDER", "byteorder")):
                if need not in cur:
                    raise TypeError(f"EdfFile: Image doesn't have {what} information")
            cur["pos"], cur["size"] = pos, size
            out.append(cur)
            cur = None
            pos += size
    return out


def _frame(path: str, index: int) -> np.ndarray:
    raw = Path(path).read_bytes()
    if raw[:2] in (b"II", b"MM"):
        raise OSError(f"'{path}' is a TIFF file behind an .edf name; use read_tiff")
    images = _images(raw)
    if index >= len(images):
        raise ValueError("EdfFile: Index out of limit")
    im = images[index]
    dims = [int(im[k]) for k in ("DIM_1", "DIM_2", "DIM_3") if k in im]
    name = im["DATATYPE"].upper()
    if name not in _TYPES:
        raise TypeError(f"unknown EdfType {name}")
    code = _TYPES[name]
    if name in ("SIGNEDLONG", "UNSIGNEDLONG") and im["size"] // max(1, int(np.prod([d if d > 0 else 1 for d in dims]))) == 8:
        code = code[0] + "8"                                         # "long" written by a 64-bit producer
    order = ">" if im["BYTEORDER"].upper() == "HIGHBYTEFIRST" else "<"
    dt = np.dtype(order + code)
    count = int(np.prod(dims))
    data = raw[im["pos"]:im["pos"] + count * dt.itemsize]
    if len(data) % dt.itemsize or len(data) < count * dt.itemsize:
        raise ValueError("buffer size must be a multiple of element size" if len(data) % dt.itemsize
                         else f"cannot reshape array of size {len(data) // dt.itemsize} into shape {tuple(dims[::-1])}")
    return np.frombuffer(data, dt).astype(dt.newbyteorder("="), copy=False).reshape(dims[::-1])


def read_edf(image_path: str | Sequence[str], *, index: int = 0, dtype: np.dtype | str = np.float32) -> np.ndarray:
    if index < 0:
        raise ValueError("index must be >= 0")

    def one(p):
        if not isinstance(p, str):
            raise TypeError("All elements of image_path must be strings")
        if not Path(p).exists():
            raise FileNotFoundError(f"EDF file not found: '{p}'")
        return np.asarray(_frame(p, index), dtype=dtype)

    if isinstance(image_path, str):
        return one(image_path)
    if not isinstance(image_path, Sequence):
        raise TypeError("image_path must be a str or a sequence of str")
    if len(image_path) == 0:
        raise ValueError("image_path sequence is empty")
    frames = []
    for p in image_path:
        arr = one(p)
        if arr.ndim != 2:
            raise ValueError(f"Expected a 2D EDF image, got shape {arr.shape} for '{p}'")
        if frames and arr.shape != frames[0].shape:
            raise ValueError(f"Inconsistent image shapes in stack: expected {frames[0].shape}, got {arr.shape} for '{p}'")
        frames.append(arr)
    return np.stack(frames, axis=0)
