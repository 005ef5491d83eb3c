"""CPU tier: the numpy oracle reproduces the outputs the REAL reference produced (tests/golden/)."""

import numpy as np
import pytest

from oracle import golden_cases as gc
from oracle import ref_numpy as orc

RT = 1e-9


def _digest_check(a, g, prefix, rtol=RT):
    a = np.asarray(a)
    cy, cx = a.shape[0] // 2, a.shape[1] // 2
    scale = float(np.max(np.abs(g[prefix + "_row"]))) + 1e-300
    for got, key in ((a[cy - 16:cy + 16, cx - 16:cx + 16], "_crop"), (a[cy, :], "_row"),
                     (a[:, cx], "_col"), (a[:8, :8], "_corner")):
        np.testing.assert_allclose(got, g[prefix + key], rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(a.sum(), g[prefix + "_sum"], rtol=1e-7, atol=1e-7 * scale)
    np.testing.assert_allclose((np.abs(a) ** 2).sum(), g[prefix + "_sumabs2"], rtol=1e-7)


@pytest.fixture(scope="module")
def frames():
    return gc.frame_cases()


@pytest.mark.parametrize("name", ["sq256", "rect128x256", "odd150x200", "u16_128", "blur256", "sq512"])
def test_signal_maps(name, frames, golden):
    g = golden("frames")
    img = frames[name]
    F, fx, fy = orc.fft2d(img)
    assert str(F.dtype) == str(g[f"{name}/fft2d_dtype"])
    _digest_check(F, g, f"{name}/fft2d", rtol=1e-6)
    np.testing.assert_array_equal(fx, g[f"{name}/fx"])
    np.testing.assert_array_equal(fy, g[f"{name}/fy"])
    P, _, _ = orc.psd2d(img)
    assert str(P.dtype) == str(g[f"{name}/psd2d_dtype"])
    _digest_check(P, g, f"{name}/psd2d", rtol=1e-6)
    Pu, _, _ = orc.psd2d(img, dx=0.5, dy=2.0)
    np.testing.assert_allclose(Pu[Pu.shape[0] // 2], g[f"{name}/psd2d_dx_row"], rtol=1e-6)
    ac, xl, yl = orc.autocorr2d(img)
    assert str(ac.dtype) == str(g[f"{name}/autocorr2d_dtype"])
    _digest_check(ac, g, f"{name}/autocorr2d")
    np.testing.assert_array_equal(xl, g[f"{name}/xlag"])
    np.testing.assert_array_equal(yl, g[f"{name}/ylag"])
    a2, _, _ = orc.autocorr2d(img, standardize=True, normalize="none")
    _digest_check(a2, g, f"{name}/autocorr2d_std_none")
    a3, _, _ = orc.autocorr2d(img, remove_mean=False, normalize="none")
    _digest_check(a3, g, f"{name}/autocorr2d_raw_none")
    xc, _, _ = orc.xcorr2d(img, np.roll(np.asarray(img), (5, -7), axis=(0, 1)))
    assert bool(np.iscomplexobj(xc)) == bool(g[f"{name}/xcorr2d_iscomplex"])
    _digest_check(np.real(xc), g, f"{name}/xcorr2d_real")
    assert tuple(np.unravel_index(int(np.argmax(np.abs(xc))), xc.shape)) == tuple(g[f"{name}/xcorr2d_argmax"])
    # quirk: peak of xcorr2d(a, roll(a, +s)) sits at lag -s
    cy, cx = xc.shape[0] // 2, xc.shape[1] // 2
    assert tuple(g[f"{name}/xcorr2d_argmax"]) == (cy - 5, cx + 7)


@pytest.mark.parametrize("name", ["sq256", "rect128x256", "odd150x200", "u16_128", "blur256", "sq512"])
def test_metrics_scalars(name, frames, golden):
    g = golden("frames")
    img = frames[name]
    for tag, kw in (("moments", {}), ("moments_nosat", {"saturation_value": None, "eps": 0.5})):
        m = orc.distribution_moments(img, **kw)
        for k, v in m.items():
            np.testing.assert_allclose(v, g[f"{name}/{tag}/{k}"], rtol=RT, equal_nan=True, err_msg=k)
    for k, v in orc.tenengrad(img).items():
        np.testing.assert_allclose(v, g[f"{name}/tenengrad/{k}"], rtol=RT, err_msg=k)
    np.testing.assert_allclose(orc.laplacian_variance(img), g[f"{name}/laplacian_variance"], rtol=RT)
    for k, v in orc.amplitude(img).items():
        np.testing.assert_allclose(v, g[f"{name}/amplitude/{k}"], rtol=RT, err_msg=k)
    np.testing.assert_allclose(orc.spectral_entropy(img), g[f"{name}/spectral_entropy"], rtol=RT)
    if min(img.shape) >= 128:
        gr = orc.grain(img)
        for k in ("lx", "ly", "leq", "r"):
            np.testing.assert_allclose(gr[k], g[f"{name}/grain/{k}"], rtol=RT, err_msg=k)
        _digest_check(gr["autocorr"], g, f"{name}/grain/autocorr")
        rad, r = orc.radial_mean_interpolated(gr["autocorr"])
        np.testing.assert_allclose(rad, g[f"{name}/grain/radial"], rtol=RT, atol=1e-12)
        np.testing.assert_array_equal(r, g[f"{name}/grain/radial_r"])
        for k, v in orc.bandwidth(img).items():
            np.testing.assert_allclose(v, g[f"{name}/bandwidth/{k}"], rtol=1e-8, err_msg=k)
        for k, v in orc.inverse_autocorr_width(img).items():
            np.testing.assert_allclose(v, g[f"{name}/inv_ac_width/{k}"], rtol=RT, err_msg=k)


def test_nonfinite_and_constant(golden):
    g = golden("frames")
    img = gc.nan_frame()
    for k, v in orc.distribution_moments(img).items():
        np.testing.assert_allclose(v, g[f"nan128/moments/{k}"], rtol=RT, equal_nan=True)
    with np.errstate(invalid="ignore"):
        for k, v in orc.tenengrad(img).items():
            np.testing.assert_allclose(v, g[f"nan128/tenengrad/{k}"], rtol=RT, equal_nan=True)
        np.testing.assert_allclose(orc.laplacian_variance(img), g["nan128/laplacian_variance"], equal_nan=True)
    const = np.full((64, 64), 7.0, dtype=np.float32)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = orc.distribution_moments(const)
    for k, v in m.items():
        np.testing.assert_allclose(v, g[f"const64/moments/{k}"], equal_nan=True)
    assert m["SNRdB"] == float("inf")


def test_tracking_matches_reference(golden):
    g = golden("tracking")
    for name, c in gc.tracking_cases().items():
        got = orc.phase_correlation(c["template"], c["image"], slices_yx=c["slices"], subpixel=c["subpixel"])
        np.testing.assert_allclose(got[:3], g[f"{name}/result"][:3], rtol=1e-6, atol=1e-6, err_msg=name)
        # noise-free frames give a background of pure rounding noise (median ~1e-6): the SNR is
        # then ill-conditioned even between two runs of the same numpy code, so it gets 1e-3.
        np.testing.assert_allclose(got[3], g[f"{name}/result"][3], rtol=1e-3, err_msg=name)


def test_tracking_quirks(golden):
    g = golden("tracking")
    # quirk 1: sub-pixel terms come back swapped: true (+0.3, -0.1) is reported as ~(-0.1, +0.3)
    dy, dx = g["sub_p3_m1_noise/result"][:2]
    assert abs(dy - (-0.1)) < abs(dy - 0.3) and abs(dx - 0.3) < abs(dx - (-0.1)) and dy < 0 < dx
    # integer rolls come back as integers to a few 1e-3
    np.testing.assert_allclose(g["roll_2_m3/result"][:2], (2, -3), atol=5e-3)
    # quirk 7: slices_yx=None with an even template raises
    a = np.zeros((64, 64), np.float32)
    with pytest.raises(ValueError):
        orc.phase_correlation(a, a, slices_yx=None)


def test_flatfield_matches_reference(golden):
    g = golden("flatfield")
    raw, flat, dark = gc.flatfield_inputs()
    for scale in ("flat_median", "flat_mean", "none"):
        out = orc.flat_field_correction(raw, flats=flat, darks=dark, scale=scale)
        assert str(out.dtype) == str(g[f"ffc/{scale}/dtype"])
        np.testing.assert_array_equal(out[:2], g[f"ffc/{scale}/frames01"])
        np.testing.assert_allclose(out.astype(np.float64).sum(), g[f"ffc/{scale}/sum"], rtol=1e-12)
    out = orc.flat_field_correction(raw, flats=flat, darks=dark, eps=50.0)
    np.testing.assert_allclose(out.astype(np.float64).sum(), g["ffc/eps50/sum"], rtol=1e-12)
    assert int((out[0] == 0).sum()) == int(g["ffc/eps50/nzero"])
    out = orc.flat_field_correction(raw[0], flats=np.stack([flat, flat + 2]), darks=np.stack([dark, dark]))
    np.testing.assert_array_equal(out, g["ffc/2d_stackflat/frame"])
    np.testing.assert_allclose(orc.flat_field_correction(raw, darks=dark).astype(np.float64).sum(),
                               g["ffc/darkonly/sum"], rtol=1e-12)
    np.testing.assert_allclose(orc.flat_field_correction(raw, flats=flat).astype(np.float64).sum(),
                               g["ffc/flatonly/sum"], rtol=1e-12)


def test_explicit_restatements_match_library_calls(frames):
    """The kernel-level definitions (stencils, bilinear polar mean, percentile) equal the library calls."""
    from scipy import ndimage
    img = frames["rect128x256"].astype(float)
    gx, gy = orc.sobel_reflect(img)
    np.testing.assert_allclose(gx, ndimage.sobel(img, axis=1, mode="reflect"), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(gy, ndimage.sobel(img, axis=0, mode="reflect"), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(orc.laplace_reflect(img), ndimage.laplace(img, mode="reflect"), rtol=1e-12, atol=1e-9)
    ac, _, _ = orc.autocorr2d(frames["sq256"])
    a, ra = orc.radial_mean_interpolated(ac)
    b, rb = orc.bilinear_polar_mean(ac)
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-13)
    np.testing.assert_array_equal(ra, rb)
    for q in (0.05, 50.0, 99.95):
        np.testing.assert_allclose(orc.percentile_linear(img, q), np.nanpercentile(img, q), rtol=1e-14)


def test_known_answers():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(64, 128)).astype(np.float32)
    # Parseval for the scaled PSD: sum(P) * nx*ny/(dx*dy) = nx*ny * sum(x^2)
    P, _, _ = orc.psd2d(a)
    np.testing.assert_allclose(P.sum() * a.size, a.size * float((a.astype(float) ** 2).sum()), rtol=1e-5)
    ac, _, _ = orc.autocorr2d(a)
    assert ac[32, 64] == 1.0 and ac.max() == 1.0
    ac0, _, _ = orc.autocorr2d(a, normalize="none")
    np.testing.assert_allclose(ac0[32, 64], a.size * a.astype(float).var(), rtol=1e-10)
    # Sobel of a linear ramp, Laplacian of a quadratic (interior)
    yy, xx = np.mgrid[0:32, 0:48].astype(float)
    gx, gy = orc.sobel_reflect(3.0 * xx + 2.0 * yy)
    assert np.allclose(gx[1:-1, 1:-1], 8 * 3.0) and np.allclose(gy[1:-1, 1:-1], 8 * 2.0)
    assert np.allclose(orc.laplace_reflect(xx ** 2 + 2 * yy ** 2)[1:-1, 1:-1], 2 + 4)
    # width of a sampled Gaussian at 1/e: 2*sqrt(2)*sigma
    x = np.arange(-200, 201, dtype=float)
    w, edge = orc.width_at_fraction(np.exp(-x ** 2 / (2 * 9.0 ** 2)))
    assert not edge and abs(w - 2 * np.sqrt(2) * 9.0) < 0.02
    w, edge = orc.width_at_fraction(np.ones(17))
    assert edge and w == 17.0


def test_tiling_executor_matches_reference(golden):
    """The oracle's tiling executor (split_edges, 3x3 / 9x9 policy, 9x9 -> 3x3 aggregation) against tiles.npz, recorded
    from the reference's speckle_stats / sharpness_stats(tiles=True) on 170/171 px tiles and 130 x 133/134 px sub-tiles."""
    g = golden("tiles")
    assert orc.split_edges(2048, 9) == [(0, 228), (228, 455), (455, 683), (683, 910), (910, 1138), (1138, 1365),
                                        (1365, 1593), (1593, 1820), (1820, 2048)]
    for name, img in gc.tile_cases().items():
        for tag, fn in (("speckle", orc.speckle_tiles), ("sharpness", orc.sharpness_tiles)):
            mode, tiles = fn(img)
            assert mode == str(g[f"{name}/{tag}/tile_mode"])
            for grp, fields in tiles.items():
                for k, v in fields.items():
                    np.testing.assert_allclose(v["mean"], g[f"{name}/{tag}/{grp}/{k}/mean"], rtol=1e-9, atol=1e-12,
                                               err_msg=f"{name} {tag} {grp}.{k} mean")
                    np.testing.assert_allclose(v["std"], g[f"{name}/{tag}/{grp}/{k}/std"], rtol=1e-7, atol=1e-12,
                                               err_msg=f"{name} {tag} {grp}.{k} std")


def test_bad_pixel_repair_matches_reference(golden):
    g = golden("flatfield_repair")
    raw, flat, dark = gc.flatfield_repair_inputs()
    np.testing.assert_array_equal(orc.flat_field_correction(raw, flat, dark, bad_pixel_removal=True), g["stack"])
    np.testing.assert_array_equal(orc.flat_field_correction(raw[1], flat, dark, scale="none", bad_pixel_removal=True), g["single"])


def test_template_matching_matches_reference(golden):
    """Oracle NCC (float64 restatement of cv2's TM_CCOEFF_NORMED) against the reference's template_matching(opencv): cv2
    evaluates the map in float32, so the map agrees to ~1e-6 and the 3x3 Taylor step to ~1e-3 px."""
    g = golden("template")
    c = gc.template_cases()["roll_25"]
    tz = orc.zscore2d(c["template"].astype(np.float32), 1e-9).astype(np.float32)
    iz = orc.zscore2d(c["image"].astype(np.float32), 1e-9).astype(np.float32)
    np.testing.assert_allclose(orc.ncc_valid(iz, tz), g["roll_25/map"], rtol=0, atol=5e-6)
    for name, c in gc.template_cases().items():
        got = orc.template_matching(c["template"], c["image"], slices_yx=c["slices"], subpixel=c["subpixel"])
        want = g[f"{name}/result"]
        np.testing.assert_allclose(got[:2], want[:2], rtol=0, atol=5e-3, err_msg=name)
        np.testing.assert_allclose(got[2], want[2], rtol=1e-5, err_msg=name + " peak")
        np.testing.assert_allclose(got[3], want[3], rtol=1e-4, err_msg=name + " snr")


@pytest.mark.parametrize("name", ["sq256", "rect128x256", "odd150x200", "u16_128", "blur256", "sq512"])
def test_eigenvalues_metric(name, frames, golden):
    """STA2 eigenvalues metric of the oracle against the reference's outputs (tests/golden/eigen.npz)."""
    g = golden("eigen")
    for k in (5, 2):
        e = orc.eigenvalues(frames[name], k=k)
        np.testing.assert_allclose([e["eigenvalues"], e["e1"], e["e2"], e["re"]], g[f"{name}/k{k}"], rtol=RT)
    with pytest.raises(ValueError):
        orc.eigenvalues(np.zeros((8, 8)))
    with pytest.raises(ValueError):
        orc.eigenvalues(frames[name], k=0)
