"""GPU tier: the drop-in metric functions and aggregators vs the goldens recorded from the reference."""

import warnings

import numpy as np
import pytest

from oracle import golden_cases as gc
from oracle import ref_numpy as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-4
POW2 = ["sq256", "rect128x256", "u16_128", "blur256", "sq512", "odd150x200"]   # odd150x200: the Bluestein path


@pytest.fixture(scope="module")
def dip():
    import barc4dip_b200 as dip
    return dip


@pytest.mark.parametrize("name", POW2)
def test_frame_metric_functions_vs_golden(dip, golden, name):
    g = golden("frames")
    img = gc.frame_cases()[name]
    m = dip.metrics.distribution_moments(img)
    for k, v in m.items():
        np.testing.assert_allclose(v, g[f"{name}/moments/{k}"], rtol=RTOL, atol=1e-12, err_msg=k)
    for k, v in dip.metrics.sharpness.tenengrad(img).items():
        np.testing.assert_allclose(v, g[f"{name}/tenengrad/{k}"], rtol=RTOL, err_msg=k)
    np.testing.assert_allclose(dip.metrics.sharpness.laplacian_variance(img), g[f"{name}/laplacian_variance"], rtol=RTOL)
    for k, v in dip.metrics.speckles.amplitude(img).items():
        np.testing.assert_allclose(v, g[f"{name}/amplitude/{k}"], rtol=RTOL, err_msg=k)
    np.testing.assert_allclose(dip.metrics.sharpness.spectral_entropy(img), g[f"{name}/spectral_entropy"], rtol=RTOL)
    gr = dip.metrics.speckles.grain(img)
    for k in ("lx", "ly", "leq", "r"):
        np.testing.assert_allclose(gr[k], g[f"{name}/grain/{k}"], rtol=RTOL, err_msg=k)
    n = max(img.shape)
    assert gr["autocorr"].shape == (n, n) and gr["autocorr"].dtype == np.float64
    np.testing.assert_allclose(gr["autocorr"][n // 2], g[f"{name}/grain/autocorr_row"], atol=1e-5)
    np.testing.assert_array_equal(gr["xlag"], np.arange(n) - n // 2)
    for k, v in dip.metrics.speckles.bandwidth(img).items():
        np.testing.assert_allclose(v, g[f"{name}/bandwidth/{k}"], rtol=RTOL, err_msg=k)
    for k, v in dip.metrics.sharpness.inverse_autocorr_width(img).items():
        np.testing.assert_allclose(v, g[f"{name}/inv_ac_width/{k}"], rtol=RTOL, err_msg=k)


def test_aggregators_vs_golden(dip, golden):
    g = golden("aggregators")
    img = gc.frame_cases()["sq256"]
    sp = dip.metrics.speckle_stats(img, tiles=False, verbose=False)
    assert sp["meta"]["kind"] == "speckles" and sp["meta"]["input_shape"] == (256, 256)
    assert sp["meta"]["requested_groups"] == ["amplitude", "bandwidth", "grain", "stats"]
    n = 0
    for key in g.files:
        if key.startswith("speckle_stats/full/"):
            _, _, grp, k = key.split("/")
            np.testing.assert_allclose(sp["full"][grp][k], g[key], rtol=RTOL, atol=1e-12, err_msg=key)
            n += 1
    assert n >= 20
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sh = dip.metrics.sharpness_stats(img, metrics=("stats", "gradient", "laplacian", "spectral", "autocorrelation"),
                                         tiles=False, verbose=False)
    for key in g.files:
        if key.startswith("sharpness_stats/full/"):
            _, _, grp, k = key.split("/")
            np.testing.assert_allclose(sh["full"][grp][k], g[key], rtol=RTOL, atol=1e-12, err_msg=key)
    stack = np.stack([gc.frame_cases()["sq256"], gc.frame_cases()["blur256"]], axis=0)
    shs = dip.metrics.sharpness_stack_stats(stack, metrics=("stats", "gradient", "laplacian"), tiles=False, verbose=False)
    assert shs["meta"]["stack_shape"] == (2, 256, 256) and shs["meta"]["n_frames"] == 2
    for key in g.files:
        if key.startswith("sharpness_stack_stats/full/"):
            _, _, grp, k = key.split("/")
            assert shs["full"][grp][k].shape == (2,)
            np.testing.assert_allclose(shs["full"][grp][k], g[key], rtol=RTOL, atol=1e-12, err_msg=key)


def test_aggregator_errors_and_schema(dip):
    from barc4dip_b200._lib import B4DUnsupported
    img = gc.frame_cases()["sq256"]
    with pytest.raises(TypeError):
        dip.metrics.speckle_stats(img.tolist())
    with pytest.raises(ValueError):
        dip.metrics.speckle_stats(img[None])
    with pytest.raises(ValueError):
        dip.metrics.speckle_stats(img, metrics="nope", tiles=False)
    with pytest.raises(ValueError):
        dip.metrics.sharpness_stats(img, tiles=False, display_origin="left")
    with pytest.warns(RuntimeWarning, match="too small for tiling"):
        small = dip.metrics.speckle_stats(img, metrics="stats", tiles=True, verbose=False)   # 256//3 < 128
    assert "tiles" not in small
    out = dip.metrics.sharpness_stats(img, tiles=False, verbose=False)
    assert set(out["full"]) == {"stats", "gradient", "laplacian", "spectral", "autocorrelation", "eigenvalues"}
    with pytest.raises(ValueError):
        dip.metrics.speckles.grain(np.ones((64, 64), np.float32))
    with pytest.raises(ValueError):
        dip.metrics.speckles.amplitude(-np.ones((128, 128), np.float32))


def test_flat_field_function_vs_golden(dip, golden):
    g = golden("flatfield")
    raw, flat, dark = gc.flatfield_inputs()
    ffc = dip.preprocessing.flat_field_correction
    for scale in ("flat_median", "none"):
        out = ffc(raw, flats=flat, darks=dark, scale=scale)
        assert out.dtype == np.float32 and out.shape == raw.shape
        np.testing.assert_array_equal(out[:2], g[f"ffc/{scale}/frames01"])
    out = ffc(raw, flats=flat, darks=dark, scale="flat_mean")
    np.testing.assert_allclose(out[:2], g["ffc/flat_mean/frames01"], rtol=3e-7)
    out = ffc(raw, flats=flat, darks=dark, eps=50.0)
    np.testing.assert_allclose(out.astype(np.float64).sum(), g["ffc/eps50/sum"], rtol=1e-12)
    assert int((out[0] == 0).sum()) == int(g["ffc/eps50/nzero"])
    out = ffc(raw[0], flats=np.stack([flat, flat + 2]), darks=np.stack([dark, dark]))
    np.testing.assert_array_equal(out, g["ffc/2d_stackflat/frame"])
    np.testing.assert_allclose(ffc(raw, darks=dark).astype(np.float64).sum(), g["ffc/darkonly/sum"], rtol=1e-12)
    np.testing.assert_allclose(ffc(raw, flats=flat).astype(np.float64).sum(), g["ffc/flatonly/sum"], rtol=1e-12)
    np.testing.assert_array_equal(ffc(raw), raw)
    with pytest.raises(ValueError):
        ffc(raw, flats=flat, scale="bogus")


def test_speckle_stack_stats_phase_tracking(dip):
    """Stack aggregator with the phase/internal tracker against the oracle's per-ROI phase correlation."""
    from barc4dip_b200 import synth
    stack, _ = synth.tracking_stack(3, 256, grain=8.0, seed=51, step_sigma=0.8)
    out = dip.metrics.speckle_stack_stats(stack, metrics=("amplitude", "stats", "grain"), tiles=False,
                                          tracking_method="phase", tracking_backend="internal", verbose=False)
    assert out["meta"]["kind"] == "speckle_stack_stats" and out["full"]["stats"]["mean"].shape == (3,)
    roi = out["meta"]["tracking"]["roi_size_yx"][0]
    step = out["meta"]["tracking"]["roi_step_yx"][0]
    g0 = orc.grain(stack[0])
    assert roi % 2 == 1 and roi >= int(np.ceil(3.0 * max(g0["lx"], g0["ly"], g0["leq"])))
    # oracle for one ROI (centre) and t = 2, absolute and incremental
    c = 128
    sl = (slice(c - roi // 2, c + roi // 2 + 1),) * 2
    want_abs = np.array([[orc.phase_correlation(stack[0][(slice(c + dy * step - roi // 2, c + dy * step + roi // 2 + 1),
                                                          slice(c + dx * step - roi // 2, c + dx * step + roi // 2 + 1))],
                                                 stack[2],
                                                 slices_yx=(slice(c + dy * step - roi // 2, c + dy * step + roi // 2 + 1),
                                                            slice(c + dx * step - roi // 2, c + dx * step + roi // 2 + 1)))[:2]
                          for dx in (-1, 0, 1)] for dy in (-1, 0, 1)])
    np.testing.assert_allclose(out["temporal"]["abs"]["dy"][2], np.float32(want_abs[..., 0].astype(np.float32).mean()), atol=0.01)
    np.testing.assert_allclose(out["temporal"]["abs"]["dx"][2], np.float32(want_abs[..., 1].astype(np.float32).mean()), atol=0.01)
    assert out["temporal"]["inc"]["dx"].dtype == np.float32 and out["temporal"]["inc"]["dx"].shape == (3,)
    for k in ("mean", "std", "skewness"):
        np.testing.assert_allclose(out["full"]["stats"][k][1], orc.distribution_moments(stack[1])[k], rtol=RTOL)


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8, np.int16, np.int32])
def test_stack_analyzer_native_integer_stacks(dtype):
    """Integer detector stacks cross PCIe in their own width and are widened on the device (b4d_cast_to_f32): the
    results are those of the same values handed over as float32, bit for bit."""
    from barc4dip_b200 import engine, synth
    from barc4dip_b200.pipeline import StackAnalyzer
    n, T = 256, 5
    stack, _ = synth.tracking_stack(T, n, grain=5.0, seed=13, integer_every=2)
    hi = {np.uint16: 60000.0, np.uint8: 250.0, np.int16: 30000.0, np.int32: 2.0e6}[dtype]
    ints = np.clip(np.rint(stack / stack.max() * hi), 0, hi).astype(dtype)
    if dtype in (np.int16, np.int32):
        ints[:, ::7, ::5] *= -1                      # signed types keep their sign
    as_f32 = ints.astype(np.float32)
    np.testing.assert_array_equal(engine.as_stack(ints).cpu().numpy(), as_f32)
    outs = []
    for data in (ints, as_f32):
        an = StackAnalyzer((n, n), reference=as_f32[0], chunk_frames=2, saturation_value=None)
        outs.append(an.run(data))
    a, b = outs
    np.testing.assert_array_equal(a["table"], b["table"])
    np.testing.assert_array_equal(a["psd"], b["psd"])
    np.testing.assert_array_equal(a["autocorr"], b["autocorr"])
    for k in ("dy", "dx", "peak", "snr"):
        np.testing.assert_array_equal(a["tracking"][k], b["tracking"][k])
    np.testing.assert_array_equal(a["amplitude"]["contrast"], b["amplitude"]["contrast"])


def _check_tiles(tiles, g, prefix, squeeze=None):
    for grp, fields in tiles.items():
        for k, v in fields.items():
            mean, std = np.asarray(v["mean"]), np.asarray(v["std"])
            wm, ws = g[f"{prefix}/{grp}/{k}/mean"], g[f"{prefix}/{grp}/{k}/std"]
            assert mean.shape == wm.shape and std.shape == ws.shape, (grp, k)
            np.testing.assert_allclose(mean, wm, rtol=RTOL, atol=1e-12, err_msg=f"{prefix} {grp}.{k} mean")
            # a spread over nine sub-tiles is a difference of nearly equal numbers: held to RTOL of the level it is about
            scale = np.maximum(np.abs(wm), 1e-300)
            assert np.all(np.isnan(std) == np.isnan(ws)), (grp, k)
            ok = ~np.isnan(ws)
            assert np.all(np.abs(std[ok] - ws[ok]) <= 3 * RTOL * scale[ok] + RTOL * np.abs(ws[ok])), f"{prefix} {grp}.{k} std"


@pytest.mark.parametrize("name", ["t3_512", "t9_1170x1200"])
def test_tiles_vs_golden(dip, golden, name):
    """speckle_stats / sharpness_stats(tiles=True) against the reference's tiling executor: 3x3 tiles of 170 / 171 px and
    9x9 sub-tiles of 130 x 133 / 134 px -- none a power of two, so every spectral metric runs on the Bluestein path."""
    g = golden("tiles")
    img = gc.tile_cases()[name]
    sp = dip.metrics.speckle_stats(img, tiles=True, verbose=False)
    assert sp["meta"]["tile_mode"] == str(g[f"{name}/speckle/tile_mode"])
    assert tuple(sp["meta"]["tile_shape_px"]) == tuple(int(v) for v in g[f"{name}/speckle/tile_shape_px"])
    assert sp["meta"]["used_subtiles"] == (name.startswith("t9"))
    assert set(sp["tiles"]) == {"amplitude", "grain", "stats", "bandwidth"}
    _check_tiles(sp["tiles"], g, f"{name}/speckle")
    sh = dip.metrics.sharpness_stats(img, tiles=True, verbose=False)
    assert set(sh["tiles"]) == set(gc.SHARPNESS_TILE_GROUPS) | {"eigenvalues"}
    sh["tiles"].pop("eigenvalues")          # pinned by eigen.npz (test_eigenvalues_metric_vs_golden)
    _check_tiles(sh["tiles"], g, f"{name}/sharpness")


def test_stack_tiles_vs_golden(dip, golden):
    """Stack aggregator: every tile leaf gets a leading T axis and each frame is display-oriented before tiling."""
    g = golden("tiles")
    img = gc.tile_cases()["t3_512"]
    stack = np.stack([img, np.ascontiguousarray(img[::-1, ::-1]) * 0.5 + 3.0], axis=0)
    shs = dip.metrics.sharpness_stack_stats(stack, metrics=("stats", "gradient", "spectral"), tiles=True, verbose=False)
    assert shs["tiles"]["stats"]["mean"]["mean"].shape == (2, 3, 3)
    _check_tiles(shs["tiles"], g, "stack/sharpness")
    # and the single-frame call on frame 1 agrees with the stack's second slice
    one = dip.metrics.sharpness_stats(stack[1], metrics=("stats", "gradient", "spectral"), tiles=True, verbose=False)
    for grp, fields in one["tiles"].items():
        for k, v in fields.items():
            np.testing.assert_allclose(v["mean"], shs["tiles"][grp][k]["mean"][1], rtol=1e-12)


def test_bad_pixel_repair_bit_exact(dip, golden):
    """flat_field_correction(bad_pixel_removal=True): bad pixels replaced by the 3x3 median (reflect border) of the frame in
    which bad pixels are zero -- bit-identical to the reference, clusters of bad pixels and borders included."""
    g = golden("flatfield_repair")
    raw, flat, dark = gc.flatfield_repair_inputs()
    ffc = dip.preprocessing.flat_field_correction
    np.testing.assert_array_equal(ffc(raw, flats=flat, darks=dark, bad_pixel_removal=True), g["stack"])
    np.testing.assert_array_equal(ffc(raw[1], flats=flat, darks=dark, scale="none", bad_pixel_removal=True), g["single"])
    assert np.count_nonzero(g["stack"] != ffc(raw, flats=flat, darks=dark)) > 0      # the repair does change pixels


def test_speckle_stack_stats_default_template_tracker(dip, golden):
    """speckle_stack_stats with the reference's default tracker (template matching on the 3x3 ROI grid, absolute and
    incremental) against goldens recorded from the reference (opencv backend): displacements within 0.01 px."""
    g = golden("template")
    stack = gc.stack_tracking_case()
    out = dip.metrics.speckle_stack_stats(stack, metrics=("stats",), tiles=False, tracking_method="template",
                                          tracking_backend="opencv", verbose=False)
    assert tuple(out["meta"]["tracking"]["roi_size_yx"]) == tuple(int(v) for v in g["stack/roi_size_yx"])
    assert tuple(out["meta"]["tracking"]["roi_step_yx"]) == tuple(int(v) for v in g["stack/roi_step_yx"])
    for mode in ("abs", "inc"):
        for k in ("dx", "dy", "r", "std_dx", "std_dy", "std_r"):
            got, want = out["temporal"][mode][k], g[f"stack/{mode}/{k}"]
            assert got.dtype == np.float32 and got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=0, atol=0.01, err_msg=f"{mode}.{k}")
    # the signature's own defaults (template / skimage) run too and give the same numbers
    dflt = dip.metrics.speckle_stack_stats(stack, metrics=("stats",), tiles=False, verbose=False)
    np.testing.assert_array_equal(dflt["temporal"]["abs"]["dx"], out["temporal"]["abs"]["dx"])


def test_two_gpu_sharding_equals_single_gpu():
    """1-GPU result == N-GPU result (tier 4): scripts/multi_gpu_check.py under torchrun with two
    ranks over NCCL -- per-frame outputs bitwise, all-reduced temporal moments within 1e-6. Needs two GPUs."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", os.path.join(root, "scripts", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multi_gpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_speckle_stats_arbitrary_frame_vs_oracle(dip):
    """A 600 x 450 frame (neither square nor a power of two; pad_to_square -> 600, Bluestein path, tiles_3x3 with 200 x 150
    tiles) against the oracle's full-frame metrics and tiling executor."""
    from barc4dip_b200 import synth
    img = synth.speckle_frame(600, 450, grain=5.0, seed=37)
    out = dip.metrics.speckle_stats(img, tiles=True, verbose=False)
    flipped = img[::-1, :]
    for grp, fn in (("amplitude", orc.amplitude), ("grain", orc.grain), ("bandwidth", orc.bandwidth),
                    ("stats", orc.distribution_moments)):
        want = fn(flipped)
        for k, v in want.items():
            if np.ndim(v) == 0:
                np.testing.assert_allclose(out["full"][grp][k], v, rtol=RTOL, atol=1e-12, err_msg=f"{grp}.{k}")
    assert out["meta"]["tile_mode"] == "tiles_3x3" and out["meta"]["tile_shape_px"] == (200, 150)
    mode, tiles = orc.speckle_tiles(img)
    assert mode == "tiles_3x3"
    for grp, fields in tiles.items():
        for k, v in fields.items():
            np.testing.assert_allclose(out["tiles"][grp][k]["mean"], v["mean"], rtol=RTOL, atol=1e-12, err_msg=f"tiles {grp}.{k}")
            assert np.all(np.isnan(out["tiles"][grp][k]["std"]))


def test_stack_tiles_chunking_and_analyzer_ragged_chunks(dip):
    """17 frames through the tiling executor (chunks of 16 frames + 1) and through StackAnalyzer with a chunk size that does
    not divide the stack: first / last frames equal the single-frame calls."""
    from barc4dip_b200 import synth
    from barc4dip_b200.pipeline import StackAnalyzer
    T, n = 17, 384                                                  # 384 // 3 = 128: tiles_3x3 of 128 px
    stack, _ = synth.tracking_stack(T, 512, grain=5.0, seed=41)
    stack = np.ascontiguousarray(stack[:, :n, :n])
    shs = dip.metrics.sharpness_stack_stats(stack, metrics=("stats", "gradient", "spectral"), tiles=True, verbose=False)
    assert shs["tiles"]["gradient"]["tenengrad"]["mean"].shape == (T, 3, 3)
    for t in (0, 15, 16):
        one = dip.metrics.sharpness_stats(stack[t], metrics=("stats", "gradient", "spectral"), tiles=True, verbose=False)
        for grp, fields in one["tiles"].items():
            for k, v in fields.items():
                np.testing.assert_allclose(v["mean"], shs["tiles"][grp][k]["mean"][t], rtol=1e-12, err_msg=f"{t} {grp}.{k}")
    sq = np.ascontiguousarray(stack[:, :256, :256])
    ref = StackAnalyzer((256, 256), reference=sq[0], chunk_frames=T).run(sq)
    for chunk in (5, 1, 32):
        got = StackAnalyzer((256, 256), reference=sq[0], chunk_frames=chunk).run(sq)
        np.testing.assert_array_equal(got["table"], ref["table"])
        np.testing.assert_array_equal(got["psd"], ref["psd"])
        np.testing.assert_array_equal(got["autocorr"], ref["autocorr"])
        for k in ("dy", "dx", "peak", "snr"):
            np.testing.assert_array_equal(got["tracking"][k], ref["tracking"][k])
        np.testing.assert_array_equal(got["amplitude"]["contrast"], ref["amplitude"]["contrast"])
    one = StackAnalyzer((256, 256), reference=sq[0]).run(sq[:1])
    np.testing.assert_array_equal(one["table"], ref["table"][:1])


def test_one_process_two_devices():
    """One process may drive several devices, each through its own context: every ABI call makes its context's device
    current for its duration (shared-memory limits are raised per device). Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from barc4dip_b200 import engine, synth
    img = synth.speckle_frame(2048, grain=6.0, seed=3)[None]
    outs = []
    for dev in (0, 1, 0):
        d = torch.from_numpy(img).to(f"cuda:{dev}")
        psd, _ = engine.psd2d(d)
        ac, g = engine.autocorr2d(d, want_grain=True)
        outs.append((psd.cpu(), ac.cpu(), g))
    assert torch.cuda.current_device() == 0
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
        np.testing.assert_array_equal(o[2], outs[0][2])


def test_stack_analyzer_arbitrary_frame_size_vs_oracle():
    """StackAnalyzer on 300 x 300 frames (not a power of two: the composed chirp-z pipeline), with a flat field fused
    into the loads: every table and map against the oracle evaluated on the corrected frames."""
    from barc4dip_b200 import synth
    from barc4dip_b200.pipeline import StackAnalyzer
    n, T = 300, 4
    stack, _ = synth.tracking_stack(T, 512, grain=5.0, seed=53, integer_every=2)
    stack = np.ascontiguousarray(stack[:, :n, :n])
    rng = np.random.default_rng(54)
    flat = (1000.0 * (1.0 + 0.1 * rng.random((n, n)))).astype(np.float32)
    dark = (100.0 + 2.0 * rng.standard_normal((n, n))).astype(np.float32)
    raw = (stack / 1000.0 * (flat - dark) + dark).astype(np.float32)
    corrected = orc.flat_field_correction(raw, flat, dark)
    an = StackAnalyzer((n, n), reference=raw[0], flats=flat, darks=dark, chunk_frames=3)
    out = an.run(raw)
    for t in range(T):
        m = orc.distribution_moments(corrected[t])
        for k in ("mean", "std", "skewness", "kurtosis"):
            np.testing.assert_allclose(out["stats"][k][t], m[k], rtol=RTOL, err_msg=k)
        np.testing.assert_allclose(out["gradient"]["tenengrad"][t], orc.tenengrad(corrected[t])["tenengrad"], rtol=RTOL)
        np.testing.assert_allclose(out["amplitude"]["contrast"][t], orc.amplitude(corrected[t])["contrast"], rtol=RTOL)
        P, _, _ = orc.psd2d(corrected[t])
        assert np.max(np.abs(out["psd"][t] - P)) <= 1e-5 * P.max()
        g = orc.grain(corrected[t])
        assert np.max(np.abs(out["autocorr"][t] - g["autocorr"])) <= 1e-5
        for k in ("lx", "ly", "leq"):
            np.testing.assert_allclose(out["grain"][k][t], g[k], rtol=RTOL, err_msg=k)
        dy, dx, peak, snr = orc.phase_correlation(corrected[0], corrected[t], slices_yx=(slice(0, n), slice(0, n)))
        np.testing.assert_allclose((out["tracking"]["dy"][t], out["tracking"]["dx"][t]), (dy, dx), atol=0.01)
        if t:
            np.testing.assert_allclose(out["tracking"]["peak"][t], peak, rtol=5e-4)


@pytest.mark.parametrize("name", ["sq256", "rect128x256", "odd150x200", "u16_128", "blur256", "sq512"])
def test_eigenvalues_metric_vs_golden(dip, golden, name):
    """STA2 eigenvalues metric (sharpness.py:752-861; library eigensolver on the device) against the reference."""
    g = golden("eigen")
    img = gc.frame_cases()[name]
    for k in (5, 2):
        e = dip.metrics.sharpness.eigenvalues(img, k=k)
        np.testing.assert_allclose([e["eigenvalues"], e["e1"], e["e2"], e["re"]], g[f"{name}/k{k}"], rtol=RTOL)
    with pytest.raises(ValueError):
        dip.metrics.sharpness.eigenvalues(np.zeros((8, 8), np.float32))
    with pytest.raises(ValueError):
        dip.metrics.sharpness.eigenvalues(img, k=0)
    if name == "sq512":
        res = dip.metrics.sharpness_stats(img, metrics="eigenvalues", tiles=True, verbose=False)
        assert res["meta"]["tile_mode"] == str(g["sq512/tile_mode"])
        np.testing.assert_allclose([res["full"]["eigenvalues"][f] for f in ("eigenvalues", "e1", "e2", "re")], g["sq512/full"], rtol=RTOL)
        for f in ("eigenvalues", "e1", "e2", "re"):
            np.testing.assert_allclose(res["tiles"]["eigenvalues"][f]["mean"], g[f"sq512/tiles/{f}/mean"], rtol=RTOL)
            np.testing.assert_allclose(res["tiles"]["eigenvalues"][f]["std"], g[f"sq512/tiles/{f}/std"], rtol=RTOL, atol=1e-12)
        st = dip.metrics.sharpness_stack_stats(np.stack([img, img[::-1]]), metrics="eigenvalues", tiles=False, verbose=False)
        np.testing.assert_allclose(st["full"]["eigenvalues"]["e1"], [g["sq512/full"][1]] * 2, rtol=RTOL)


def test_fused_blocks_equal_composed_blocks_and_use_one_forward_fft(dip):
    """speckle_stats / sharpness_stats on square power-of-two frames take every FFT-based group from ONE fused pass
    (b4d_stack_pipeline_ref with spectral sums): same numbers as the per-metric blocks, one forward transform per frame
    in the launch list (reference: metrics/speckles.py:740-796, metrics/sharpness.py:581-629)."""
    from barc4dip_b200 import engine, stack as blocks, synth
    from barc4dip_b200._lib import get_context
    frames = np.stack([synth.speckle_frame(512, grain=5.0, seed=s) for s in (81, 82, 83)])
    d = engine.as_stack(frames)
    fb = blocks.FusedBlocks(d, saturation_value=65535.0, eps=1e-6, keep_map=True)
    table = engine.frame_reductions(d, saturation_value=65535.0, eps=1e-6)
    np.testing.assert_array_equal(fb.table, table)
    for got, want in ((fb.amplitude(), blocks.amplitude_block(d, table)), (fb.grain(), blocks.grain_block(d, table=table)),
                      (fb.bandwidth(), blocks.bandwidth_block(d, table=table)), (fb.entropy(), blocks.spectral_entropy_block(d)),
                      (fb.inverse_autocorr(), blocks.inverse_autocorr_block(d, table=table))):
        assert set(got) == set(want)
        for k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=2e-5, err_msg=k)
    # launch list: one forward row pass per call, whatever the number of FFT-based groups
    ctx = get_context()
    for fn, kw in ((dip.metrics.speckle_stats, {}), (dip.metrics.sharpness_stats, {"metrics": ("stats", "gradient", "laplacian", "spectral", "autocorrelation")})):
        fn(frames[0], tiles=False, verbose=False, **kw)            # warm (tables, scratch)
        ctx.profile_begin()
        fn(frames[0], tiles=False, verbose=False, **kw)
        prof = ctx.profile_end()
        assert prof["rows_fwd"][1] == 1, prof
