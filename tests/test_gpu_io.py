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
