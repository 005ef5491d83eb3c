"""
Golden fixtures of the file readers  --  TEST INFRASTRUCTURE ONLY (run in the build container, never on the product path).

EDF: small files are assembled here by hand from the format's public description (ASCII header `{ key = value ; ... }`
padded to a block multiple, binary data right behind it) and read by the REAL reference (`barc4dip.io.edf.read_edf`,
io/edf.py:18-91, over its parser io/uti_EdfFile.py). File bytes and the reference's arrays go to tests/golden/edf.npz;
tests/test_cpu_io.py replays the bytes through barc4dip_b200.io.edf.read_edf. (The reference's own EDF *writer* no longer
runs on numpy 2 -- ndarray.tostring -- which is why the files are assembled here.)
TIFF: files written by PIL, read by the reference's read_tiff (io/tiff.py:19-70) -> tests/golden/tiff.npz.

    python oracle/make_golden_io.py
"""

from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.load_reference import REFERENCE_SRC, reference_available  # noqa: E402

EDF_TYPES = {"uint8": "UnsignedByte", "int8": "SignedByte", "uint16": "UnsignedShort", "int16": "SignedShort",
             "uint32": "UnsignedInteger", "int32": "SignedInteger", "uint64": "Unsigned64", "int64": "Signed64",
             "float32": "FloatValue", "float64": "DoubleValue"}


def edf_image(a: np.ndarray, *, type_name: str | None = None, big_endian: bool = False, block: int = 512, extra: str = "",
              number: int = 1, eol: str = "\n", lower_keys: bool = False) -> bytes:
    """One EDF image (header block + data) holding array `a` ((Dim_2, Dim_1) or (Dim_3, Dim_2, Dim_1))."""
    dims = a.shape[::-1]
    keys = ["HeaderID", "Image", "ByteOrder", "DataType"] + [f"Dim_{i + 1}" for i in range(len(dims))] + ["Size"]
    vals = [f"EH:{number:06d}:000000:000000", str(number), "HighByteFirst" if big_endian else "LowByteFirst",
            type_name or EDF_TYPES[str(a.dtype)]] + [str(d) for d in dims] + [str(a.nbytes)]
    if lower_keys:
        keys = [k.lower() if k.startswith(("Dim", "Size")) else k for k in keys]
    h = "{" + eol + "".join(f"{k} = {v} ;{eol}" for k, v in zip(keys, vals)) + extra
    pad = -(len(h) + 1 + len(eol)) % block
    h += " " * pad + "}" + eol
    data = a.astype(a.dtype.newbyteorder(">" if big_endian else "<")).tobytes()
    return h.encode("latin-1") + data


def edf_cases() -> dict[str, tuple[bytes, int]]:
    """name -> (file bytes, frame index to read)."""
    rng = np.random.default_rng(12)
    out = {}
    for name in EDF_TYPES:
        dt = np.dtype(name)
        a = rng.standard_normal((6, 9)).astype(dt) * 100 if dt.kind == "f" else \
            rng.integers(max(np.iinfo(dt).min, -2 ** 40), min(np.iinfo(dt).max, 2 ** 40), size=(6, 9)).astype(dt)
        out[f"le_{name}"] = (edf_image(a), 0)
    a16 = rng.integers(0, 65535, size=(7, 5)).astype(np.uint16)
    f32 = (rng.standard_normal((7, 5)) * 1e3).astype(np.float32)
    out["be_uint16"] = (edf_image(a16, big_endian=True), 0)
    out["be_float32_block1024"] = (edf_image(f32, big_endian=True, block=1024), 0)
    out["extra_keys"] = (edf_image(a16, extra="Title = speckle scan = 3 ;\ncount_time = 0.05 ;\nmotor_pos = 1.5 2.5 ;\n"), 0)
    out["crlf_lowercase_keys"] = (edf_image(a16, eol="\r\n", lower_keys=True), 0)
    out["long_as_32bit"] = (edf_image(a16.astype(np.int32), type_name="SignedLong"), 0)
    out["unsignedlong_as_64bit"] = (edf_image(a16.astype(np.uint64), type_name="UnsignedLong"), 0)
    out["float_alias"] = (edf_image(f32, type_name="Float"), 0)
    two = edf_image(a16, number=1) + edf_image(f32, number=2, block=1024)
    out["two_images_first"] = (two, 0)
    out["two_images_second"] = (two, 1)
    out["three_dims"] = (edf_image(rng.integers(0, 255, size=(3, 4, 5)).astype(np.uint8)), 0)
    return out


def main():
    if not reference_available():
        raise SystemExit("needs /root/reference")
    if "barc4dip" not in sys.modules:
        stub = types.ModuleType("barc4dip")
        stub.__path__ = [REFERENCE_SRC]
        sys.modules["barc4dip"] = stub
        iostub = types.ModuleType("barc4dip.io")              # io/__init__ imports h5py: skip it, load the modules
        iostub.__path__ = [os.path.join(REFERENCE_SRC, "io")]
        sys.modules["barc4dip.io"] = iostub
    ref_edf = importlib.import_module("barc4dip.io.edf")
    ref_tiff = importlib.import_module("barc4dip.io.tiff")
    gold = os.path.join(os.path.dirname(HERE), "tests", "golden")
    save = {}
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for name, (raw, index) in edf_cases().items():
            p = os.path.join(tmp, name + ".edf")
            with open(p, "wb") as fh:
                fh.write(raw)
            save[f"{name}/file"] = np.frombuffer(raw, np.uint8)
            save[f"{name}/index"] = np.int64(index)
            save[f"{name}/float32"] = ref_edf.read_edf(p, index=index)
            save[f"{name}/float64"] = ref_edf.read_edf(p, index=index, dtype=np.float64)
            if name in ("be_uint16", "extra_keys", "crlf_lowercase_keys"):
                paths.append(p)
        save["sequence/names"] = np.array([os.path.basename(p)[:-4] for p in paths])
        save["sequence/float32"] = ref_edf.read_edf(paths)
        np.savez_compressed(os.path.join(gold, "edf.npz"), **save)

        from PIL import Image
        rng = np.random.default_rng(13)
        tsave = {}
        tpaths = []
        for name, arr in (("u16", rng.integers(0, 65535, size=(11, 13)).astype(np.uint16)),
                          ("u8", rng.integers(0, 255, size=(11, 13)).astype(np.uint8)),
                          ("f32", rng.standard_normal((11, 13)).astype(np.float32)),
                          ("i32", rng.integers(-2 ** 31, 2 ** 31 - 1, size=(11, 13)).astype(np.int32)),
                          ("u16_b", rng.integers(0, 65535, size=(11, 13)).astype(np.uint16))):
            p = os.path.join(tmp, name + ".tif")
            Image.fromarray(arr).save(p)
            with open(p, "rb") as fh:
                tsave[f"{name}/file"] = np.frombuffer(fh.read(), np.uint8)
            tsave[f"{name}/array"] = ref_tiff.read_tiff(p)
            if name.startswith("u16"):
                tpaths.append(p)
        tsave["sequence/names"] = np.array([os.path.basename(p)[:-4] for p in tpaths])
        tsave["sequence/array"] = ref_tiff.read_tiff(tpaths)
        np.savez_compressed(os.path.join(gold, "tiff.npz"), **tsave)
    print("wrote", os.path.join(gold, "edf.npz"), "and tiff.npz")


if __name__ == "__main__":
    main()
