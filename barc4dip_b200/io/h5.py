"""
read_h5 / save_h5 with the reference's signatures, dataset location and error behaviour (io/h5.py:17-142, :145-212).

Backend: h5py when it can be imported (the reference's own), otherwise the self-contained codec in `hdf5.py`. The files
are the ESRF layout either way: one dataset at /entry_0000/measurement/data, written chunked with gzip level 4 and the
NX_class attributes of the two groups (io/h5.py:196-210).
"""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import numpy as np

from . import hdf5

DATASET_PATH = "entry_0000/measurement/data"

try:                                                   # pragma: no cover - not installed in this image
    import h5py as _h5py
except ImportError:
    _h5py = None


def backend() -> str:
    return "h5py" if _h5py is not None else "builtin"


def open_dataset(path):
    """(file object to close, dataset) of the stack in `path`; both backends give shape / dtype / slicing on the dataset."""
    f = _h5py.File(path, "r") if _h5py is not None else hdf5.H5File(path)
    try:
        if DATASET_PATH not in f:
            raise KeyError(f"Dataset not found: '{DATASET_PATH}' in '{path}'")
        return f, f[DATASET_PATH]
    except BaseException:
        f.close()
        raise


def _read_one(p, image_number):
    if not isinstance(p, str):
        raise TypeError("All elements of image_path must be strings")
    if not Path(p).exists():
        raise FileNotFoundError(f"HDF5 file not found: '{p}'")
    try:
        f, dset = open_dataset(p)
        try:
            if image_number is None:
                arr = dset[()]
            else:
                if dset.ndim != 3:
                    raise ValueError("image_number is only valid for 3D datasets (N, H, W); "
                                     f"got shape {dset.shape} in '{p}'")
                n_frames = int(dset.shape[0])
                idx = int(image_number)
                if idx < 0:
                    idx += n_frames
                if idx < 0 or idx >= n_frames:
                    raise ValueError(f"image_number={image_number} out of bounds for dataset with {n_frames} frames in '{p}'")
                arr = dset[idx, :, :]
        finally:
            f.close()
    except OSError as e:
        raise OSError(f"Failed to read HDF5 file: '{p}'") from e
    arr = np.asarray(arr)
    if arr.ndim not in (2, 3):
        raise ValueError(f"Expected 2D or 3D dataset at '{DATASET_PATH}', got shape {arr.shape} in '{p}'")
    return arr


def read_h5(image_path: str | Sequence[str], *, image_number: int = None) -> np.ndarray:
    """One file -> its dataset ((H, W), (N, H, W), or frame `image_number` of a stack); a sequence of files -> one
    (N, H, W) stack (2-D files stacked, 3-D files concatenated along axis 0)."""
    if isinstance(image_path, str):
        return _read_one(image_path, image_number)
    if image_number is not None:
        raise ValueError("image_number is only supported when image_path is a single file (str)")
    if not isinstance(image_path, Sequence):
        raise TypeError("image_path must be a str or a sequence of str")
    if len(image_path) == 0:
        raise ValueError("image_path sequence is empty")
    arrays = [_read_one(p, None) for p in image_path]
    ndims = {a.ndim for a in arrays}
    if ndims == {2}:
        for p, a in zip(image_path, arrays):
            if a.shape != arrays[0].shape:
                raise ValueError(f"Inconsistent image shapes in stack: expected {arrays[0].shape}, got {a.shape} for '{p}'")
        return np.stack(arrays, axis=0)
    if ndims == {3}:
        for p, a in zip(image_path, arrays):
            if a.shape[1:] != arrays[0].shape[1:]:
                raise ValueError(f"Inconsistent stack shapes: expected (*, {arrays[0].shape[1:]}), got {a.shape} for '{p}'")
        return np.concatenate(arrays, axis=0)
    raise ValueError(f"Mixed dataset dimensionality across files: ndims={sorted(ndims)}")


def save_h5(data: np.ndarray, output_path: str | Path) -> None:
    """(H, W) image or (N, H, W) stack -> a new HDF5 file (never overwrites; a suffix other than .h5 / .hdf5 becomes .h5)."""
    if not isinstance(data, np.ndarray):
        raise TypeError("data must be a numpy.ndarray")
    if data.ndim not in (2, 3):
        raise ValueError(f"data must be 2D or 3D, got ndim={data.ndim}")
    out = Path(output_path)
    if out.name == "":
        raise ValueError("output_path must include a filename")
    if not out.parent.exists():
        raise OSError(f"Invalid path: directory does not exist: {out.parent}")
    if not out.parent.is_dir():
        raise OSError(f"Invalid path: not a directory: {out.parent}")
    if out.suffix.lower() not in {".h5", ".hdf5"}:
        out = out.with_suffix(".h5")
    if out.exists():
        raise OSError(f"Refusing to overwrite existing file: {out}")
    try:
        if _h5py is not None:                          # pragma: no cover
            with _h5py.File(out, "x") as f:
                entry = f.require_group("entry_0000")
                meas = entry.require_group("measurement")
                entry.attrs.setdefault("NX_class", "NXentry")
                meas.attrs.setdefault("NX_class", "NXcollection")
                meas.create_dataset("data", data=data, compression="gzip", compression_opts=4, chunks=True)
        else:
            hdf5.write_stack(out, data, dataset_path=DATASET_PATH, compression=4, chunks=True,
                             group_attrs={"entry_0000": {"NX_class": "NXentry"}, "measurement": {"NX_class": "NXcollection"}})
    except OSError as e:
        raise OSError(f"Failed to write HDF5 file: {out}") from e
