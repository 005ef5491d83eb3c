"""
Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference)
on the seeded inputs of ``oracle/golden_cases.py``.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

The outputs are small (scalars, 1-D cuts, 32x32 crops, and a few float32 maps at 128-256 px)
and are committed; the inputs are regenerated from seeds by the tests.
Recorded with numpy 2.3.5 / scipy 1.18.1 (the reference pins no versions).
"""

from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import golden_cases as gc          # noqa: E402
from oracle.load_reference import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def crop_center(a, h=16):
    cy, cx = a.shape[0] // 2, a.shape[1] // 2
    return np.array(a[cy - h:cy + h, cx - h:cx + h])


def map_digest(a, prefix, store):
    """Compact, tight pin of a 2-D map: centre crop, centre cuts, corner, global sums."""
    a = np.asarray(a)
    cy, cx = a.shape[0] // 2, a.shape[1] // 2
    store[prefix + "_crop"] = crop_center(a)
    store[prefix + "_row"] = np.array(a[cy, :])
    store[prefix + "_col"] = np.array(a[:, cx])
    store[prefix + "_corner"] = np.array(a[:8, :8])
    store[prefix + "_sum"] = np.asarray(a.sum())
    store[prefix + "_sumabs2"] = np.asarray((np.abs(a) ** 2).sum())


def signal1d_golden(sig, versions):
    """signal1d.npz: the reference's 1-D helpers (signal/fft.py:31-196, signal/corr.py:45-166)."""
    store = {"versions": versions}
    for name, a in gc.signal1d_cases().items():
        F, fx = sig.fft.fft1d(a, dx=0.5)
        store[f"{name}/fft"], store[f"{name}/fx"] = F, fx
        store[f"{name}/ifft"] = sig.fft.ifft1d(F)
        P, _ = sig.fft.psd1d(a, dx=0.5)
        store[f"{name}/psd"] = P
        b = np.roll(a, 7)
        c, lag = sig.corr.xcorr1d(a, b, dx=0.5)
        store[f"{name}/xcorr"], store[f"{name}/lag"] = np.real(c), lag
        c2, _ = sig.corr.autocorr1d(a)
        store[f"{name}/autocorr"] = np.real(c2)
        c3, _ = sig.corr.autocorr1d(a, remove_mean=False, normalize="none")
        store[f"{name}/autocorr_raw"] = np.real(c3)
    np.savez_compressed(os.path.join(OUT, "signal1d.npz"), **store)
    print("signal1d.npz", os.path.getsize(os.path.join(OUT, "signal1d.npz")) // 1024, "KiB")


def eigen_golden(met, versions):
    """eigen.npz: the STA2 eigenvalues metric (metrics/sharpness.py:752-861) on the frame cases, and through
    sharpness_stats(metrics="eigenvalues") with tiles."""
    store = {"versions": versions}
    for name, img in gc.frame_cases().items():
        for k in (5, 2):
            e = met.sharpness.eigenvalues(img, k=k)
            store[f"{name}/k{k}"] = np.array([e["eigenvalues"], e["e1"], e["e2"], e["re"]])
    img = gc.frame_cases()["sq512"]
    res = met.sharpness.sharpness_stats(img, metrics="eigenvalues", tiles=True, verbose=False)
    store["sq512/full"] = np.array([res["full"]["eigenvalues"][f] for f in ("eigenvalues", "e1", "e2", "re")])
    for f in ("eigenvalues", "e1", "e2", "re"):
        store[f"sq512/tiles/{f}/mean"] = np.asarray(res["tiles"]["eigenvalues"][f]["mean"])
        store[f"sq512/tiles/{f}/std"] = np.asarray(res["tiles"]["eigenvalues"][f]["std"])
    store["sq512/tile_mode"] = np.array(res["meta"]["tile_mode"])
    np.savez_compressed(os.path.join(OUT, "eigen.npz"), **store)
    print("eigen.npz", os.path.getsize(os.path.join(OUT, "eigen.npz")) // 1024, "KiB")


def template_golden(sig, versions):
    """template.npz: template_matching(backend="opencv") of the reference (cv2.matchTemplate TM_CCOEFF_NORMED)."""
    import cv2
    store = {"versions": versions, "cv2": np.array(cv2.__version__)}
    for name, c in gc.template_cases().items():
        res = sig.template_matching(c["template"], c["image"], slices_yx=c["slices"], backend="opencv", subpixel=c["subpixel"])
        store[f"{name}/result"] = np.asarray(res, dtype=np.float64)
        res2 = sig.track_translation(c["template"], c["image"], slices_yx=c["slices"], method="template", backend="opencv",
                                     subpixel=c["subpixel"])
        assert tuple(res2) == tuple(res)
    c = gc.template_cases()["roll_25"]
    tz = sig.tracking._zscore2d(c["template"].astype(np.float32), eps=1e-9).astype(np.float32)
    iz = sig.tracking._zscore2d(c["image"].astype(np.float32), eps=1e-9).astype(np.float32)
    store["roll_25/map"] = cv2.matchTemplate(iz, tz, method=cv2.TM_CCOEFF_NORMED)
    # the stack aggregator with its default tracker ("template"; opencv backend, skimage is not installed here)
    from oracle.load_reference import load_reference
    met = load_reference().metrics
    st = gc.stack_tracking_case()
    res = met.speckle_stack_stats(st, metrics=("stats",), tiles=False, tracking_method="template", tracking_backend="opencv",
                                  parallel=False, verbose=False)
    for mode in ("abs", "inc"):
        for k, v in res["temporal"][mode].items():
            store[f"stack/{mode}/{k}"] = np.asarray(v)
    store["stack/roi_size_yx"] = np.asarray(res["meta"]["tracking"]["roi_size_yx"])
    store["stack/roi_step_yx"] = np.asarray(res["meta"]["tracking"]["roi_step_yx"])
    np.savez_compressed(os.path.join(OUT, "template.npz"), **store)
    print("template.npz", os.path.getsize(os.path.join(OUT, "template.npz")) // 1024, "KiB")


def repair_golden(pre, versions):
    """flatfield_repair.npz: flat_field_correction(bad_pixel_removal=True) of the reference (3x3 median repair)."""
    raw, flat, dark = gc.flatfield_repair_inputs()
    store = {"versions": versions}
    out = pre.flat_field_correction(raw, flats=flat, darks=dark, bad_pixel_removal=True)
    store["stack"] = out
    store["single"] = pre.flat_field_correction(raw[1], flats=flat, darks=dark, scale="none", bad_pixel_removal=True)
    np.savez_compressed(os.path.join(OUT, "flatfield_repair.npz"), **store)
    print("flatfield_repair.npz", os.path.getsize(os.path.join(OUT, "flatfield_repair.npz")) // 1024, "KiB")


def tiles_golden(met, versions):
    """tiles.npz: the reference's tiling executor on the frames of golden_cases.tile_cases()."""
    store = {"versions": versions}
    for name, img in gc.tile_cases().items():
        sp = met.speckle_stats(img, tiles=True, verbose=False)
        sh = met.sharpness_stats(img, metrics=gc.SHARPNESS_TILE_GROUPS, tiles=True, verbose=False)
        for tag, res in (("speckle", sp), ("sharpness", sh)):
            store[f"{name}/{tag}/tile_mode"] = np.array(res["meta"]["tile_mode"])
            store[f"{name}/{tag}/tile_shape_px"] = np.asarray(res["meta"]["tile_shape_px"])
            for grp, fields in res["tiles"].items():
                for k, v in fields.items():
                    store[f"{name}/{tag}/{grp}/{k}/mean"] = np.asarray(v["mean"], dtype=np.float64)
                    store[f"{name}/{tag}/{grp}/{k}/std"] = np.asarray(v["std"], dtype=np.float64)
    # stack variant: leading T axis, per-frame display orientation
    img = gc.tile_cases()["t3_512"]
    stack = np.stack([img, np.ascontiguousarray(img[::-1, ::-1]) * 0.5 + 3.0], axis=0)
    shs = met.sharpness_stack_stats(stack, metrics=("stats", "gradient", "spectral"), tiles=True, verbose=False, parallel=False)
    for grp, fields in shs["tiles"].items():
        for k, v in fields.items():
            store[f"stack/sharpness/{grp}/{k}/mean"] = np.asarray(v["mean"], dtype=np.float64)
            store[f"stack/sharpness/{grp}/{k}/std"] = np.asarray(v["std"], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "tiles.npz"), **store)


def schema_golden(met):
    """schema.json: key tree / kinds / dtypes / shapes of the reference's aggregator dicts, plus the number-masked text
    of the reference's own logbook_report on them (report/markdown.py:37)."""
    import importlib
    import json
    from oracle import schema as sc
    rep = importlib.import_module("barc4dip.report.markdown")
    out = {}
    for name, call in sc.schema_calls(met).items():
        d = call()
        entry = {"tree": sc.schema_tree(d)}
        try:
            entry["markdown"] = {f"complete={c}": sc.markdown_skeleton(rep.logbook_report(d, complete=c, notes=False)) for c in (False, True)}
        except ValueError as e:                        # the stack variants have no formatter in the reference
            entry["markdown_error"] = str(e)
        out[name] = entry
        print(name, len(entry["tree"]), "leaves", "markdown" if "markdown" in entry else entry["markdown_error"])
    with open(os.path.join(OUT, "schema.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


def main():
    ref = load_reference()
    sig, met, pre = ref.signal, ref.metrics, ref.preprocessing_normalize
    os.makedirs(OUT, exist_ok=True)
    import scipy
    versions = np.array([np.__version__, scipy.__version__])
    if "--only-eigen" in sys.argv:
        eigen_golden(met, versions)
        return
    if "--only-1d" in sys.argv:
        signal1d_golden(sig, versions)
        return
    if "--only-template" in sys.argv:
        template_golden(sig, versions)
        return
    if "--only-repair" in sys.argv:
        repair_golden(pre, versions)
        return
    if "--only-schema" in sys.argv:
        schema_golden(met)
        return
    if "--only-tiles" in sys.argv:
        tiles_golden(met, versions)
        print("tiles.npz", os.path.getsize(os.path.join(OUT, "tiles.npz")) // 1024, "KiB")
        return

    # ---------------- per-frame functions ----------------
    store = {"versions": versions}
    for name, img in gc.frame_cases().items():
        F, fx, fy = sig.fft2d(img)
        map_digest(F, f"{name}/fft2d", store)
        store[f"{name}/fft2d_dtype"] = np.array(str(F.dtype))
        store[f"{name}/fx"] = fx
        store[f"{name}/fy"] = fy
        P, _, _ = sig.psd2d(img)
        map_digest(P, f"{name}/psd2d", store)
        store[f"{name}/psd2d_dtype"] = np.array(str(P.dtype))
        if img.shape[0] <= 256 and img.dtype == np.float32:
            store[f"{name}/psd2d_full"] = np.asarray(P, dtype=np.float32)
        Pu, _, _ = sig.psd2d(img, dx=0.5, dy=2.0, scale=True)
        store[f"{name}/psd2d_dx_row"] = np.array(Pu[Pu.shape[0] // 2, :])
        ac, xl, yl = sig.autocorr2d(img)
        map_digest(ac, f"{name}/autocorr2d", store)
        store[f"{name}/autocorr2d_dtype"] = np.array(str(ac.dtype))
        if img.shape[0] <= 256 and img.dtype == np.float32:
            store[f"{name}/autocorr2d_full32"] = np.asarray(ac, dtype=np.float32)
        store[f"{name}/xlag"] = xl
        store[f"{name}/ylag"] = yl
        acs, _, _ = sig.autocorr2d(img, remove_mean=True, standardize=True, normalize="none")
        map_digest(acs, f"{name}/autocorr2d_std_none", store)
        acr, _, _ = sig.autocorr2d(img, remove_mean=False, standardize=False, normalize="none")
        map_digest(acr, f"{name}/autocorr2d_raw_none", store)
        other = np.roll(np.asarray(img), (5, -7), axis=(0, 1))
        xc, _, _ = sig.xcorr2d(img, other)
        store[f"{name}/xcorr2d_iscomplex"] = np.array(np.iscomplexobj(xc))
        map_digest(np.real(xc), f"{name}/xcorr2d_real", store)
        store[f"{name}/xcorr2d_argmax"] = np.array(np.unravel_index(int(np.argmax(np.abs(xc))), xc.shape))

        for k, v in met.distribution_moments(img).items():
            store[f"{name}/moments/{k}"] = np.asarray(v)
        for k, v in met.distribution_moments(img, saturation_value=None, eps=0.5).items():
            store[f"{name}/moments_nosat/{k}"] = np.asarray(v)
        for k, v in met.sharpness.tenengrad(img).items():
            store[f"{name}/tenengrad/{k}"] = np.asarray(v)
        store[f"{name}/laplacian_variance"] = np.asarray(met.sharpness.laplacian_variance(img))
        for k, v in met.speckles.amplitude(img).items():
            store[f"{name}/amplitude/{k}"] = np.asarray(v)
        store[f"{name}/spectral_entropy"] = np.asarray(met.sharpness.spectral_entropy(img))
        if min(img.shape) >= 128:
            g = met.speckles.grain(img)
            for k in ("lx", "ly", "leq", "r"):
                store[f"{name}/grain/{k}"] = np.asarray(g[k])
            map_digest(g["autocorr"], f"{name}/grain/autocorr", store)
            rad, r = ref.maths.radial.radial_mean_interpolated(g["autocorr"])
            store[f"{name}/grain/radial"] = rad
            store[f"{name}/grain/radial_r"] = r
            for k, v in met.speckles.bandwidth(img).items():
                store[f"{name}/bandwidth/{k}"] = np.asarray(v)
            for k, v in met.sharpness.inverse_autocorr_width(img).items():
                store[f"{name}/inv_ac_width/{k}"] = np.asarray(v)

    nanimg = gc.nan_frame()
    for k, v in met.distribution_moments(nanimg).items():
        store[f"nan128/moments/{k}"] = np.asarray(v)
    for k, v in met.sharpness.tenengrad(nanimg).items():
        store[f"nan128/tenengrad/{k}"] = np.asarray(v)
    store["nan128/laplacian_variance"] = np.asarray(met.sharpness.laplacian_variance(nanimg))
    const = np.full((64, 64), 7.0, dtype=np.float32)
    for k, v in met.distribution_moments(const).items():
        store[f"const64/moments/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "frames.npz"), **store)

    # ---------------- aggregator schema (full-frame, tiles off) ----------------
    store = {"versions": versions}
    img = gc.frame_cases()["sq256"]
    sp = met.speckle_stats(img, tiles=False, verbose=False)
    for grp, d in sp["full"].items():
        for k, v in d.items():
            if np.ndim(v) == 0:
                store[f"speckle_stats/full/{grp}/{k}"] = np.asarray(v)
    sh = met.sharpness_stats(img, metrics=("stats", "gradient", "laplacian", "spectral", "autocorrelation"),
                             tiles=False, verbose=False)
    for grp, d in sh["full"].items():
        for k, v in d.items():
            store[f"sharpness_stats/full/{grp}/{k}"] = np.asarray(v)
    stack = np.stack([gc.frame_cases()["sq256"], gc.frame_cases()["blur256"]], axis=0)
    shs = met.sharpness_stack_stats(stack, metrics=("stats", "gradient", "laplacian"), tiles=False,
                                    verbose=False, parallel=False)
    for grp, d in shs["full"].items():
        for k, v in d.items():
            store[f"sharpness_stack_stats/full/{grp}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "aggregators.npz"), **store)

    # ---------------- tracking ----------------
    store = {"versions": versions}
    for name, c in gc.tracking_cases().items():
        res = sig.phase_correlation(c["template"], c["image"], slices_yx=c["slices"],
                                    backend="internal", subpixel=c["subpixel"], eps=1e-9)
        store[f"{name}/result"] = np.asarray(res, dtype=np.float64)
        res2 = sig.track_translation(c["template"], c["image"], slices_yx=c["slices"], method="phase",
                                     backend="internal", subpixel=c["subpixel"])
        assert tuple(res2) == tuple(res)
    np.savez_compressed(os.path.join(OUT, "tracking.npz"), **store)

    # ---------------- flat field ----------------
    store = {"versions": versions}
    raw, flat, dark = gc.flatfield_inputs()
    for scale in ("flat_median", "flat_mean", "none"):
        out = pre.flat_field_correction(raw, flats=flat, darks=dark, scale=scale)
        store[f"ffc/{scale}/frames01"] = out[:2]
        store[f"ffc/{scale}/sum"] = np.asarray(out.astype(np.float64).sum())
        store[f"ffc/{scale}/dtype"] = np.array(str(out.dtype))
    out = pre.flat_field_correction(raw, flats=flat, darks=dark, eps=50.0)
    store["ffc/eps50/sum"] = np.asarray(out.astype(np.float64).sum())
    store["ffc/eps50/nzero"] = np.asarray(int((out[0] == 0).sum()))
    out = pre.flat_field_correction(raw[0], flats=np.stack([flat, flat + 2]), darks=np.stack([dark, dark]))
    store["ffc/2d_stackflat/frame"] = out
    out = pre.flat_field_correction(raw, darks=dark)
    store["ffc/darkonly/sum"] = np.asarray(out.astype(np.float64).sum())
    out = pre.flat_field_correction(raw, flats=flat)
    store["ffc/flatonly/sum"] = np.asarray(out.astype(np.float64).sum())
    np.savez_compressed(os.path.join(OUT, "flatfield.npz"), **store)

    tiles_golden(met, versions)
    repair_golden(pre, versions)
    template_golden(sig, versions)
    signal1d_golden(sig, versions)
    eigen_golden(met, versions)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
