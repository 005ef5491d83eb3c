"""GPU tier: frame reductions, selection, flat field and temporal moments vs the oracle and the goldens."""

import numpy as np
import pytest

from oracle import golden_cases as gc
from oracle import ref_numpy as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # BASELINE.json north_star: metrics within 1e-4 relative


@pytest.fixture(scope="module")
def eng():
    from barc4dip_b200 import engine
    return engine


def _moments(eng, img, **kw):
    from barc4dip_b200.metrics.statistics import moments_from_row
    sat = kw.get("saturation_value", 65535.0)
    tab = eng.frame_reductions(eng.as_stack(img), **kw)
    return moments_from_row(tab[0], sat), tab[0]


@pytest.mark.parametrize("name", ["sq256", "rect128x256", "odd150x200", "u16_128", "blur256", "sq512"])
def test_frame_reductions_vs_golden(eng, golden, name):
    from barc4dip_b200._lib import FR
    g = golden("frames")
    img = gc.frame_cases()[name]
    m, row = _moments(eng, img)
    for k, v in m.items():
        np.testing.assert_allclose(v, g[f"{name}/moments/{k}"], rtol=RTOL, atol=1e-12, err_msg=k)
    m2, _ = _moments(eng, img, saturation_value=None, eps=0.5)
    for k, v in m2.items():
        np.testing.assert_allclose(v, g[f"{name}/moments_nosat/{k}"], rtol=RTOL, atol=1e-12, equal_nan=True, err_msg=k)
    n = row[FR["count"]]
    ex, ey = row[FR["sgx2"]] / n, row[FR["sgy2"]] / n
    np.testing.assert_allclose(ex, g[f"{name}/tenengrad/ex"], rtol=RTOL)
    np.testing.assert_allclose(ey, g[f"{name}/tenengrad/ey"], rtol=RTOL)
    lapvar = row[FR["slap2"]] / n - (row[FR["slap"]] / n) ** 2
    np.testing.assert_allclose(lapvar, g[f"{name}/laplacian_variance"], rtol=RTOL)


def test_frame_reductions_nonfinite_and_constant(eng, golden):
    from barc4dip_b200._lib import FR
    g = golden("frames")
    img = gc.nan_frame()
    m, row = _moments(eng, img)
    for k, v in m.items():
        np.testing.assert_allclose(v, g[f"nan128/moments/{k}"], rtol=RTOL, atol=1e-12, err_msg=k)
    assert row[FR["nnan"]] == 1 and row[FR["npix"]] - row[FR["count"]] == 3
    n = row[FR["count"]]
    np.testing.assert_allclose(row[FR["sgx2"]] / n, g["nan128/tenengrad/ex"], rtol=RTOL, equal_nan=True)
    const = np.full((64, 64), 7.0, dtype=np.float32)
    m, _ = _moments(eng, const)
    assert m["std"] == 0.0 and m["mean"] == 7.0 and m["SNRdB"] == float("inf")


def test_frame_reductions_stack_and_fused_flatfield(eng):
    import torch
    raw, flat, dark = gc.flatfield_inputs()
    den = flat - dark
    eps = 1e-6 * float(np.median(den))
    s = float(np.median(den[den > eps]))
    d_raw, d_flat, d_dark = (eng.as_stack(raw), torch.from_numpy(flat).cuda(), torch.from_numpy(dark).cuda())
    gain = eng.flat_gain(d_flat, d_dark, eps=eps, scale_value=s)
    tab = eng.frame_reductions(d_raw, gain=gain, dark=d_dark)
    corr = orc.flat_field_correction(raw, flats=flat, darks=dark)
    from barc4dip_b200.metrics.statistics import moments_from_row
    for t in range(raw.shape[0]):
        want = orc.distribution_moments(corr[t])
        got = moments_from_row(tab[t], 65535.0)
        for k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=RTOL, atol=1e-9, err_msg=f"frame {t} {k}")
        ten = orc.tenengrad(corr[t])
        np.testing.assert_allclose(tab[t][7] / tab[t][0], ten["ex"], rtol=RTOL)


@pytest.mark.parametrize("shape", [(1, 37, 53), (3, 64, 130), (2, 257, 129), (2, 1024, 1024)])
def test_frame_reductions_ragged_shapes_vs_oracle(eng, shape):
    rng = np.random.default_rng(7)
    stack = (rng.gamma(2.0, 500.0, size=shape)).astype(np.float32)
    tab = eng.frame_reductions(eng.as_stack(stack))
    from barc4dip_b200.metrics.statistics import moments_from_row
    for t in range(shape[0]):
        want = orc.distribution_moments(stack[t])
        got = moments_from_row(tab[t], 65535.0)
        for k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=RTOL, atol=1e-12, err_msg=k)
        ten = orc.tenengrad(stack[t])
        np.testing.assert_allclose(tab[t][7] / tab[t][0], ten["ex"], rtol=RTOL)
        np.testing.assert_allclose(tab[t][8] / tab[t][0], ten["ey"], rtol=RTOL)
        lv = tab[t][10] / tab[t][0] - (tab[t][9] / tab[t][0]) ** 2
        np.testing.assert_allclose(lv, orc.laplacian_variance(stack[t]), rtol=RTOL)


def test_low_contrast_frame_keeps_precision(eng):
    """mean >> std: the shifted accumulation must not lose the higher moments (quirk list: float64 reference)."""
    rng = np.random.default_rng(3)
    img = (10000.0 + rng.normal(0, 2.0, size=(512, 512)) + (rng.random((512, 512)) < 0.01) * 30.0).astype(np.float32)
    want = orc.distribution_moments(img)
    from barc4dip_b200.metrics.statistics import moments_from_row
    got = moments_from_row(eng.frame_reductions(eng.as_stack(img))[0], 65535.0)
    for k in ("mean", "std", "variance", "skewness", "kurtosis"):
        np.testing.assert_allclose(got[k], want[k], rtol=RTOL, err_msg=k)


@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 150, 200), (3, 512, 512)])
def test_select_quantiles_exact(eng, shape):
    rng = np.random.default_rng(11)
    stack = rng.exponential(1000.0, size=shape).astype(np.float32)
    stack[0, 0, :5] = np.nan
    stack[0, 1, 1] = -3.5
    qs = [0.0005, 0.9995]
    vals, nv = eng.select_quantiles(eng.as_stack(stack), qs)
    for t in range(shape[0]):
        f = stack[t].ravel()
        s = np.sort(f[~np.isnan(f)])
        assert nv[t] == s.size
        for i, q in enumerate(qs):
            h = eng.virtual_index(s.size, q)
            lo = int(np.floor(h))
            assert vals[t, 2 * i] == s[lo] and vals[t, 2 * i + 1] == s[min(lo + 1, s.size - 1)]
            got = eng.quantile_from_bracket(vals[t, 2 * i], vals[t, 2 * i + 1], s.size, q)
            np.testing.assert_allclose(got, np.nanpercentile(stack[t].astype(np.float64), 100 * q), rtol=1e-12)
    # median of |x| (even count): mean of the two middle values
    vals, nv = eng.select_quantiles(eng.as_stack(-stack[-1:]), [0.5], use_abs=True)
    s = np.abs(stack[-1]).ravel()
    s = np.sort(s[~np.isnan(s)])
    n = s.size
    assert nv[0] == n
    assert vals[0, 0] == s[(n - 1) // 2] and vals[0, 1] == s[min((n - 1) // 2 + 1, n - 1)]


def test_flat_field_bit_exact(eng, golden):
    import torch
    g = golden("flatfield")
    raw, flat, dark = gc.flatfield_inputs()
    den = flat - dark
    eps = np.float32(1e-6) * np.median(den)
    for scale in ("flat_median", "flat_mean", "none"):
        bad = den <= eps
        s = {"flat_median": np.median(den[~bad]), "flat_mean": np.mean(den[~bad]), "none": np.float32(1)}[scale]
        out = eng.flat_field(eng.as_stack(raw), torch.from_numpy(flat).cuda(), torch.from_numpy(dark).cuda(),
                             eps=float(eps), scale_value=float(s), apply_scale=scale != "none").cpu().numpy()
        np.testing.assert_array_equal(out[:2], g[f"ffc/{scale}/frames01"])
        np.testing.assert_allclose(out.astype(np.float64).sum(), g[f"ffc/{scale}/sum"], rtol=1e-12)


@pytest.mark.parametrize("with_ff", [False, True])
def test_temporal_moments_vs_oracle(eng, with_ff):
    import torch
    raw, flat, dark = gc.temporal_inputs()
    if with_ff:
        den = flat - dark
        eps = 1e-6 * float(np.median(den))
        s = float(np.median(den[den > eps]))
        gain = eng.flat_gain(torch.from_numpy(flat).cuda(), torch.from_numpy(dark).cuda(), eps=eps, scale_value=s)
        got = eng.temporal_moments(eng.as_stack(raw), gain=gain, dark=torch.from_numpy(dark).cuda())
        ref_in = orc.flat_field_correction(raw, flats=flat, darks=dark)
    else:
        got = eng.temporal_moments(eng.as_stack(raw))
        ref_in = raw
    want = orc.temporal_moments(ref_in)
    ok = want["std"] > 0          # dead pixels are constant 0 -> skew/kurt NaN in both
    for k in ("mean", "std", "variance"):
        np.testing.assert_allclose(got[k], want[k], rtol=RTOL, atol=1e-9, err_msg=k)
    for k in ("skewness", "kurtosis"):
        np.testing.assert_allclose(got[k][ok], want[k][ok], rtol=RTOL, atol=1e-6, err_msg=k)


def test_temporal_moments_chunked_equals_single(eng):
    rng = np.random.default_rng(5)
    stack = rng.exponential(1000.0, size=(70, 96, 128)).astype(np.float32)
    d = eng.as_stack(stack)
    one = eng.temporal_moments(d)
    acc = eng.TemporalAccumulator(96, 128, device=0)
    acc.pilot(d)
    for a, b in ((0, 13), (13, 50), (50, 70)):
        acc.update(d[a:b])
    many = acc.finalize()
    want = orc.temporal_moments(stack)
    for k in want:
        np.testing.assert_allclose(many[k], one[k], rtol=1e-5, atol=1e-6)   # fp32 batches fall on other frames
        np.testing.assert_allclose(many[k], want[k], rtol=RTOL, atol=1e-6)


def _tail_brackets(frame, q):
    f = frame.ravel().astype(np.float32)
    s = np.sort(f[~np.isnan(f)])
    n = s.size
    h = n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0
    lo = int(np.clip(np.floor(h), 0, n - 1))
    return s[lo], s[min(lo + 1, n - 1)], n


@pytest.mark.parametrize("shape", [(3, 512, 512), (2, 1024, 1024), (1, 256, 2048)])
def test_fused_tail_percentiles_exact(eng, shape):
    """The order statistics collected inside the reduction pass are the exact np.nanpercentile neighbours."""
    rng = np.random.default_rng(21)
    stack = rng.exponential(1000.0, size=shape).astype(np.float32)
    stack[0, 0, :5] = np.nan
    stack[0, 3, 7] = -np.inf
    stack[0, 5, 9] = np.inf
    stack[0, 1, 1] = -3.5
    stack[-1, 10, :40] = 0.0                       # zero candidates (slow path of the screening)
    stack[-1, 11, :40] = 70000.0                   # saturated pixels
    q_lo, q_hi = 0.0005, 0.9995
    tab, quant, nv = eng.frame_reductions_tails(eng.as_stack(stack), q_lo, q_hi)
    tab, quant, nv = tab.cpu().numpy(), quant.cpu().numpy(), nv.cpu().numpy()
    ref_tab = eng.frame_reductions(eng.as_stack(stack))
    np.testing.assert_allclose(tab, ref_tab, rtol=0, atol=0, equal_nan=True)      # same kernel with and without the tails
    for t in range(shape[0]):
        a0, a1, n = _tail_brackets(stack[t], q_lo)
        b0, b1, _ = _tail_brackets(stack[t], q_hi)
        assert nv[t] == n
        assert (quant[t, 0], quant[t, 1], quant[t, 2], quant[t, 3]) == (a0, a1, b0, b1)
        lo = eng.quantile_from_bracket(quant[t, 0], quant[t, 1], int(nv[t]), q_lo)
        hi = eng.quantile_from_bracket(quant[t, 2], quant[t, 3], int(nv[t]), q_hi)
        want = np.nanpercentile(stack[t].astype(np.float64), [100 * q_lo, 100 * q_hi])
        np.testing.assert_allclose([lo, hi], want, rtol=1e-6)


def test_fused_tail_percentiles_flag_and_fallback(eng):
    """A frame whose tails are not small is flagged, never wrong. Frame 0 fools the probe: its 16384 strided sample
    positions hold large values, every other pixel a small distinct one, so nearly the whole frame lies below the lower
    threshold and the candidate list overflows. A constant frame (frame 1) is one run of ties and needs no fallback."""
    rng = np.random.default_rng(2)
    n = 512 * 512
    f0 = rng.uniform(0.0, 1.0, size=n).astype(np.float32)
    f0[(np.arange(16384, dtype=np.int64) * n) // 16384] = (100.0 + rng.uniform(0.0, 1.0, size=16384)).astype(np.float32)
    stack = np.stack([f0.reshape(512, 512), np.full((512, 512), 5.0, np.float32),
                      rng.normal(100.0, 5.0, size=(512, 512)).astype(np.float32)])
    d = eng.as_stack(stack)
    _, quant, nv = eng.frame_reductions_tails(d, 0.0005, 0.9995)
    assert int(nv[0]) == -1 and int(nv[1]) == n and int(nv[2]) == n
    assert tuple(quant[1].cpu().numpy()) == (5.0, 5.0, 5.0, 5.0)
    eng.resolve_tail_quantiles(d, quant, nv, 0.0005, 0.9995)
    a0, a1, _ = _tail_brackets(stack[0], 0.0005)
    b0, b1, _ = _tail_brackets(stack[0], 0.9995)
    assert int(nv[0]) == n and tuple(quant[0].cpu().numpy()) == (a0, a1, b0, b1)


@pytest.mark.parametrize("zeros,sat", [(0.3, 0.01), (0.004, 0.0), (0.0, 0.2), (0.0004, 0.0004)])
def test_fused_tail_percentiles_integer_frames_with_ties(eng, zeros, sat):
    """Detector frames are integer valued and tie by the thousand at 0 (masked / dark regions) and at the saturation
    value: the runs of ties at the thresholds are counted, not listed, so such frames resolve inside the reduction pass
    (no fallback) and give np.sort's order statistics exactly."""
    rng = np.random.default_rng(int(zeros * 1e4) + int(sat * 1e4) * 7 + 1)
    stack = np.round(rng.exponential(1000.0, size=(2, 1024, 1024))).astype(np.float32)
    u = rng.uniform(size=stack.shape)
    stack[u < zeros] = 0.0
    stack[u > 1.0 - sat] = 65535.0
    q_lo, q_hi = 0.0005, 0.9995
    _, quant, nv = eng.frame_reductions_tails(eng.as_stack(stack), q_lo, q_hi)
    quant, nv = quant.cpu().numpy(), nv.cpu().numpy()
    for t in range(2):
        a0, a1, n = _tail_brackets(stack[t], q_lo)
        b0, b1, _ = _tail_brackets(stack[t], q_hi)
        assert nv[t] == n, "frame was not resolved inside the reduction pass"
        assert (quant[t, 0], quant[t, 1], quant[t, 2], quant[t, 3]) == (a0, a1, b0, b1)


def test_fused_tail_percentiles_with_flat_field(eng):
    import torch
    from barc4dip_b200 import synth
    raw, flat, dark = synth.flatfield_case(2, 512, seed=33, dead_frac=2e-4)
    den = flat - dark
    eps = 1e-6 * float(np.median(den))
    s = float(np.median(den[den > eps]))
    d_raw, d_flat, d_dark = (eng.as_stack(raw), torch.from_numpy(flat).cuda(), torch.from_numpy(dark).cuda())
    gain = eng.flat_gain(d_flat, d_dark, eps=eps, scale_value=s)
    _, quant, nv = eng.frame_reductions_tails(d_raw, 0.0005, 0.9995, gain=gain, dark=d_dark)
    corr = eng.flat_field(d_raw, d_flat, d_dark, eps=eps, scale_value=s, apply_scale=True)
    want, nv2 = eng.select_quantiles(corr, [0.0005, 0.9995], return_device=True)
    ok = (nv >= 0)
    assert bool(ok.any())
    # the fused loader evaluates (raw - dark) * gain, the materialising kernel (raw - dark) / den * s: 1-ulp differences
    np.testing.assert_allclose(quant[ok].cpu().numpy(), want[ok].cpu().numpy(), rtol=1e-6)
    assert bool((nv[ok] == nv2[ok]).all())
