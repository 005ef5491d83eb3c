"""
Stack ingestion: the reference's HDF5 entry points (io/h5.py, io/rw.py) and a streaming reader that
feeds the fused stack pipeline from a file without ever holding the stack in host memory.

    barc4dip_b200.io.read_image / write_image     io/rw.py:64, :151 (HDF5, TIFF and EDF in, HDF5 out)
    barc4dip_b200.io.h5.read_h5 / save_h5         io/h5.py:17, :145
    barc4dip_b200.io.stream.analyze_h5_stack      file -> pinned staging -> StackAnalyzer, decode overlapped with the GPU
    barc4dip_b200.io.hdf5                          the self-contained HDF5 codec behind them when h5py is absent
"""

from . import edf, h5, hdf5, stream, tiff  # noqa: F401
from .rw import read_image, write_image  # noqa: F401

__all__ = ["edf", "h5", "hdf5", "stream", "tiff", "read_image", "write_image"]
