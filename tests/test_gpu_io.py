"""
File -> GPU ingestion (io.stream): a stack read block by block from HDF5 through pinned staging gives the results of the
same frames handed to StackAnalyzer.run in memory, bit for bit (reference path: read_image -> per-frame loop,
io/rw.py:129, metrics/speckles.py:300-325).
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b, path=""):
    assert type(a) is type(b), path
    if isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    else:
        np.testing.assert_array_equal(a, b, err_msg=path)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32, np.float64])
def test_streamed_file_equals_in_memory_stack(tmp_path, dtype):
    from barc4dip_b200 import synth
    from barc4dip_b200.io import h5 as h5io
    from barc4dip_b200.io.stream import analyze_h5_stack
    from barc4dip_b200.pipeline import StackAnalyzer
    n, T = 256, 11
    stack, _ = synth.tracking_stack(T, n, grain=5.0, seed=21, integer_every=3)
    data = np.clip(np.rint(stack / stack.max() * 60000.0), 0, 60000).astype(dtype) if dtype == np.uint16 else stack.astype(dtype)
    p = str(tmp_path / "scan.h5")
    h5io.save_h5(data, p)
    in_memory = data if dtype != np.float64 else data.astype(np.float32)          # float64 files are staged as float32
    want = StackAnalyzer((n, n), reference=in_memory[0], chunk_frames=4).run(in_memory)
    got = analyze_h5_stack(p, block_frames=4, chunk_frames=4, want_maps=True)      # blocks of 4 + 4 + 3 frames
    _same(got, want)
    # tables only (the default), odd block size, explicit reference and a frame range: one rank's share of the stack
    part = analyze_h5_stack(p, frames=(3, 10), block_frames=5, reference=in_memory[0])
    assert "psd" not in part
    for grp in ("stats", "gradient", "grain", "amplitude", "tracking"):
        for k, v in part[grp].items():
            np.testing.assert_array_equal(v, want[grp][k][3:10], err_msg=f"{grp}/{k}")
    np.testing.assert_array_equal(part["table"], want["table"][3:10])


def test_streamed_analysis_reuses_an_analyzer_and_keeps_maps_on_device(tmp_path):
    import torch
    from barc4dip_b200 import synth
    from barc4dip_b200.io import h5 as h5io
    from barc4dip_b200.io.stream import analyze_h5_stack
    from barc4dip_b200.pipeline import StackAnalyzer
    n, T = 128, 6
    stack = synth.speckle_stack(T, n, grain=4.0).astype(np.float32)
    p = str(tmp_path / "scan.h5")
    h5io.save_h5(stack, p)
    an = StackAnalyzer((n, n), reference=stack[0], chunk_frames=2)
    want = an.run(stack, keep_maps_on_device=True)
    got = analyze_h5_stack(p, analyzer=an, block_frames=4, keep_maps_on_device=True)
    assert isinstance(got["psd"], torch.Tensor) and got["psd"].shape == (T, n, n)
    assert torch.equal(got["psd"], want["psd"]) and torch.equal(got["autocorr"], want["autocorr"])
    np.testing.assert_array_equal(got["table"], want["table"])


def _chunk_major(frames: np.ndarray, chunks, shuffle: bool) -> np.ndarray:
    """The bytes HDF5 holds after inflation: whole chunks in [frame block][tile row][tile column] order, edge chunks
    padded, each chunk byte-shuffled when the shuffle filter is on."""
    c0, cy, cx = chunks
    T, ny, nx = frames.shape
    es = frames.dtype.itemsize
    out = []
    for f in range(0, T, c0):
        for y in range(0, ny, cy):
            for x in range(0, nx, cx):
                blk = np.zeros(chunks, frames.dtype)
                part = frames[f:f + c0, y:y + cy, x:x + cx]
                blk[:part.shape[0], :part.shape[1], :part.shape[2]] = part
                raw = np.frombuffer(blk.tobytes(), np.uint8)
                out.append(np.ascontiguousarray(raw.reshape(-1, es).T).reshape(-1) if shuffle else raw)
    return np.concatenate(out)


@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int16", "int32", "uint32", "float32"])
@pytest.mark.parametrize("shape,chunks", [((5, 64, 96), (1, 16, 96)), ((5, 64, 96), (2, 24, 32)), ((3, 37, 50), (2, 8, 7)),
                                          ((4, 40, 44), (3, 40, 12))])
@pytest.mark.parametrize("shuffle", [False, True])
def test_unchunk_kernel_against_numpy(dtype, shape, chunks, shuffle):
    """b4d_unchunk_to_f32: chunk tiles (ragged edges, several frames per chunk) + byte shuffle + widening, bit-exact,
    for whole stacks and for a frame range that starts inside a chunk."""
    import torch
    from barc4dip_b200._lib import UNCHUNK_CODES, get_context, ptr
    rng = np.random.default_rng(7)
    if dtype == "float32":
        frames = rng.standard_normal(shape).astype(np.float32)
    else:
        info = np.iinfo(dtype)
        frames = rng.integers(info.min, info.max, size=shape, endpoint=True, dtype=dtype)
    ctx = get_context()
    dev = torch.from_numpy(_chunk_major(frames, chunks, shuffle)).cuda()
    for lo, hi in [(0, shape[0]), (1, shape[0]), (shape[0] - 1, shape[0])]:
        fb0 = lo // chunks[0]
        per_fb = -(-shape[1] // chunks[1]) * -(-shape[2] // chunks[2]) * int(np.prod(chunks)) * frames.dtype.itemsize
        out = torch.full((hi - lo,) + shape[1:], -1.0, dtype=torch.float32, device="cuda")
        src = dev[fb0 * per_fb:]
        ctx.check(ctx.lib.b4d_unchunk_to_f32(ctx.handle, ptr(src), UNCHUNK_CODES[dtype], int(shuffle), hi - lo, shape[1], shape[2],
                                             *chunks, lo - fb0 * chunks[0], ptr(out)), "b4d_unchunk_to_f32")
        np.testing.assert_array_equal(out.cpu().numpy(), frames[lo:hi].astype(np.float32))
    with pytest.raises(ValueError):
        ctx.check(ctx.lib.b4d_unchunk_to_f32(ctx.handle, ptr(dev), 1, 0, 1, 8, 8, 2, 8, 8, 2, ptr(out)), "first >= c0")


def _needs_engine():
    from barc4dip_b200._lib import inflate_caps
    mask, max_bytes = inflate_caps()
    if not mask & 1:
        pytest.skip("no hardware deflate engine on this device / driver")
    return max_bytes


@pytest.mark.parametrize("dtype,chunks,shuffle", [("uint16", (1, 64, 256), False), ("uint16", (2, 100, 128), True),
                                                  ("float32", (1, 256, 256), True), ("uint8", (3, 50, 60), False)])
def test_device_inflater_delivers_the_file_s_frames(tmp_path, dtype, chunks, shuffle):
    """Compressed chunks -> PCIe -> decompression engine -> unchunk kernel == the array that was written."""
    from barc4dip_b200.io import hdf5
    from barc4dip_b200.io.stream import DeviceInflater
    assert _needs_engine() >= 1 << 20
    T, n = 9, 256
    rng = np.random.default_rng(3)
    base = rng.gamma(1.0, 500.0, size=(T, n, n))
    data = base.astype(dtype) if dtype == "float32" else np.clip(base, 0, np.iinfo(dtype).max).astype(dtype)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, data, chunks=chunks, shuffle=shuffle)
    with hdf5.H5File(p) as f:
        dset = f["entry_0000/measurement/data"]
        assert DeviceInflater.unsupported(dset) is None
        for frames, block in [(None, 4), ((2, 9), 3), ((1, 2), 32)]:
            lo0 = 0 if frames is None else frames[0]
            src = DeviceInflater(dset, frames=frames, block_frames=block)
            sizes = [hi - lo for lo, hi, _ in src.blocks]
            assert max(sizes) <= block and sum(sizes) == (T if frames is None else frames[1] - frames[0])
            got = [(lo, blk.cpu().numpy()) for lo, blk in src]
            assert got[0][0] == lo0 and [len(g) for _, g in got] == sizes
            hi0 = T if frames is None else frames[1]
            np.testing.assert_array_equal(np.concatenate([g for _, g in got]), data[lo0:hi0].astype(np.float32))


def test_device_inflate_equals_host_inflate_end_to_end(tmp_path):
    from barc4dip_b200 import synth
    from barc4dip_b200.io import h5 as h5io, hdf5
    from barc4dip_b200.io.stream import DeviceInflater, analyze_h5_stack
    _needs_engine()
    n, T = 256, 10
    stack, _ = synth.tracking_stack(T, n, grain=5.0, seed=4)
    data = np.clip(np.rint(stack / stack.max() * 60000.0), 0, 60000).astype(np.uint16)
    p = str(tmp_path / "scan.h5")
    h5io.save_h5(data, p)
    host = analyze_h5_stack(p, inflate="host", block_frames=4, want_maps=True)
    dev = analyze_h5_stack(p, inflate="device", block_frames=4, want_maps=True)
    _same(dev, host)
    _same(analyze_h5_stack(p, block_frames=3, want_maps=True), host)           # "auto" takes the device path here
    part = analyze_h5_stack(p, inflate="device", frames=(3, 8), reference=data[0])
    np.testing.assert_array_equal(part["table"], host["table"][3:8])
    # files the engine cannot take fall back to the host under "auto" and are refused under "device"
    q = str(tmp_path / "plain.h5")
    hdf5.write_stack(q, data, compression=None, chunks=(1, 64, 256))
    with hdf5.H5File(q) as f:
        assert "deflate" in DeviceInflater.unsupported(f["entry_0000/measurement/data"])
    np.testing.assert_array_equal(analyze_h5_stack(q, block_frames=5)["table"], host["table"])
    with pytest.raises(OSError, match="cannot be inflated on the device"):
        analyze_h5_stack(q, inflate="device")
    with pytest.raises(ValueError):
        analyze_h5_stack(p, inflate="gpu")
