"""Pins the CPU oracle (oracle/oracle_np.py) to the reference: the shipped iodide
known-answer set and the outputs of the unmodified reference on seeded synthetic
inputs (tests/golden, made by oracle/make_golden.py).  CPU only."""
import csv
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_params
from oracle import oracle_np as orc


def _csv(path):
    with open(path) as fh:
        rd = list(csv.reader(fh))
    return rd[0], rd[1:]


def _cols(path):
    hdr, rows = _csv(path)
    return hdr, {h: [r[k] for r in rows] for k, h in enumerate(hdr)}


# ---------------------------------------------------------------- shipped known-answer set
def test_iodide_golden_rates_legacy_simpson():
    hdr, rows = _csv(os.path.join(GOLDEN, "reference_iodide", "I-ADF-efg-I-acfs.csv"))
    a = np.array([[float(x) for x in r[:7]] for r in rows])
    time, f = a[:, 1], a[:, 2:7].T
    r = orc.qr_from_acf(time, f, "127I", simpson_mode="avg")
    _, rr = _csv(os.path.join(GOLDEN, "reference_iodide", "I-ADF-efg-I-relax.csv"))
    want = [float(x) for x in rr[0][2:7]]
    got = [r["T1"], r["T2"], r["Tiso"], r["tau_c"], r["V0"]]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # f_{2,-m} == f_{2,m} (SURVEY §4 self-consistency)
    np.testing.assert_allclose(f[0], f[4], rtol=0, atol=1e-15)
    np.testing.assert_allclose(f[1], f[3], rtol=0, atol=1e-15)


def test_simpson_matches_scipy_both_parities():
    sp = pytest.importorskip("scipy.integrate")
    rng = np.random.default_rng(0)
    for n in (2, 3, 4, 5, 10, 11, 316, 317):
        x = np.cumsum(rng.uniform(0.5, 1.5, n))
        y = rng.normal(size=n)
        assert abs(orc.simpson(y, x) - sp.simpson(y, x=x)) <= 1e-12 * max(1, abs(sp.simpson(y, x=x)))
    x = np.arange(316) * 0.029
    y = np.exp(-x)
    assert abs(orc.simpson(y, x, "avg") - (1 - np.exp(-x[-1]))) < 1e-5


def test_wk_equals_direct_and_interpolates():
    rng = np.random.default_rng(1)
    z = rng.normal(size=37) + 1j * rng.normal(size=37)
    np.testing.assert_allclose(orc.wiener_khinchin(z), orc.acf_direct(z), rtol=0, atol=1e-13)
    x = np.array([np.nan, 1, np.nan, 3, np.nan])
    np.testing.assert_array_equal(orc.interpolate_nan(x)[1:], [1, 2, 3, 3])
    assert np.isnan(orc.interpolate_nan(x)[0])


# ---------------------------------------------------------------- QR end to end
@pytest.mark.parametrize("case", ["qr_small", "qr_odd", "qr_nan"])
def test_qr_pipeline_vs_reference(case):
    rows = orc.read_efg_csv(os.path.join(GOLDEN, case, "efg.csv"))
    res, acfs = orc.qr_pipeline(rows, "127I")
    hdr, ref = _csv(os.path.join(GOLDEN, case, "ref", "efg-I-relax.csv"))
    for r in ref:
        if r[0] in ("mean", "err"):
            continue
        got = res[int(r[0])]
        np.testing.assert_allclose([got["T1"], got["T2"], got["Tiso"], got["tau_c"], got["V0"]],
                                   [float(x) for x in r[2:7]], rtol=1e-11)
    hdr, ref = _csv(os.path.join(GOLDEN, case, "ref", "efg-I-acfs.csv"))
    a = np.array([[float(x) for x in r[:7]] for r in ref])
    mine = np.concatenate([np.vstack([acfs[t][0], acfs[t][1], acfs[t][2]]).T for t in sorted(acfs)])
    assert mine.shape == a.shape
    np.testing.assert_allclose(mine, a, rtol=0, atol=1e-12)


# ---------------------------------------------------------------- DD end to end
def _run_dd(case, traj="01"):
    P = load_params(case)
    PD, DD = P.ParseDynamics, P.Dipolar
    pos, sym, frames = orc.read_xyz(os.path.join(GOLDEN, case, traj, "traj.xyz"), PD.nat, PD.start_prod,
                                    PD.end_prod, PD.sample_freq, PD.md_print_freq)
    return orc.dd_pipeline(pos, sym, frames, PD.celldm, PD.timestep, identifier=DD.identifier,
                           pair_type=getattr(DD, "pair_type", "all"), rmax=getattr(DD, "rmax", 15),
                           analyte_label=getattr(DD, "analyte_label", None),
                           analyte_species=getattr(DD, "analyte_species", "1H"))


@pytest.mark.parametrize("case,trajs", [("dd_gapless_all", ["01"]), ("dd_gaps_inter", ["01"]),
                                         ("dd_intra_sampled", ["01", "02"]), ("dd_analyte_label", ["01"]),
                                         ("dd_hetero_13C", ["01"]), ("dd_dynamic_all", ["01"]),
                                         ("dd_dynamic_inter", ["01"]), ("dd_dynamic_intra", ["01"])])
def test_dd_pipeline_vs_reference(case, trajs):
    _, res = _csv(os.path.join(GOLDEN, case, "ref", "DD-results.csv"))
    for k, traj in enumerate(trajs):
        out = _run_dd(case, traj)
        hdr, rows = _csv(os.path.join(GOLDEN, case, "ref", "%s__avg-acf.csv" % traj))
        a = np.array([[float(x) for x in r] for r in rows])
        assert a.shape[0] == len(out["time"])
        np.testing.assert_allclose(out["time"], a[:, 1], rtol=1e-14)
        for c in range(7):
            ref = a[:, 2 + c]
            # G_r of rigid intramolecular pairs is rounding noise (r^-3 - mean = 0): judged on the scale of the
            # r^-3-weighted columns
            scale = np.max(np.abs(a[:, 6:9])) if c == 3 else np.max(np.abs(ref))
            assert np.max(np.abs(out["G"][c] - ref)) <= 1e-11 * scale
        want = [float(x) for x in res[k][1:4]]
        np.testing.assert_allclose([out["rate"], out["tau_c"], out["F0"]], want, rtol=1e-10)


# ---------------------------------------------------------------- SR end to end
@pytest.mark.parametrize("case", ["sr_methane", "sr_water", "sr_methane_ar", "sr_methyl", "sr_methyl_global", "sr_ring",
                                  "sr_tinker_vel"])
def test_sr_pipeline_vs_reference(case):
    P = load_params(case)
    PD, SR = P.ParseDynamics, P.SpinRotation
    sel = (PD.nat, PD.start_prod, PD.end_prod, PD.sample_freq, PD.md_print_freq)
    vel = None
    if PD.MD_format == "Tinker":   # explicit velocities; two waters touch in one frame -> they are dropped
        pos, sym, frames = orc.read_tinker(os.path.join(GOLDEN, case, "01", "water.arc"), *sel)
        vel, _, _ = orc.read_tinker(os.path.join(GOLDEN, case, "01", "water.vel"), *sel)
    else:
        pos, sym, frames = orc.read_xyz(os.path.join(GOLDEN, case, "01", "traj.xyz"), *sel)
    out = orc.sr_pipeline(pos, sym, frames, PD.celldm, PD.timestep, SR.mol_type, SR.nmol, SR.C_SR,
                          identifier=getattr(SR, "identifier", None), methyl_indeces=getattr(SR, "methyl_indeces", None),
                          global_momentum=getattr(SR, "global_momentum", False), vel=vel)
    ref = os.path.join(GOLDEN, case, "ref")
    hdr, rows = _csv(os.path.join(ref, "01__J_cart.csv"))
    a = np.array([[float(x) for x in r[1:]] for r in rows])
    tab = out["table"]
    assert a.shape[0] == len(tab["frame"])
    np.testing.assert_array_equal(a[:, 3].astype(int), tab["frame"])
    np.testing.assert_array_equal(a[:, 4].astype(int), tab["molecule"])
    np.testing.assert_array_equal(a[:, 5].astype(int), tab["label"])
    assert np.max(np.abs(a[:, :3] - tab["J"])) <= 1e-10 * np.max(np.abs(a[:, :3]))
    hdr, rows = _csv(os.path.join(ref, "01__molax.csv"))
    a = np.array([[float(x) for x in r[1:10]] for r in rows])
    np.testing.assert_allclose(tab["ax"], a, rtol=0, atol=1e-12)
    hdr, rows = _csv(os.path.join(ref, "01__ang_vel_molax.csv"))
    a = np.array([[float(x) for x in r[1:4]] for r in rows])
    assert np.max(np.abs(a - tab["w"])) <= 1e-10 * np.max(np.abs(a))
    hdr, rows = _csv(os.path.join(ref, "01__Jacf.csv"))
    a = np.array([[float(x) for x in r] for r in rows])
    np.testing.assert_array_equal(a[:, 0].astype(int), out["frames"])
    for c in range(3):
        assert np.max(np.abs(a[:, 1 + c] - out["acf_mean"][c])) <= 1e-11 * np.max(np.abs(a[:, 1 + c]))
    np.testing.assert_allclose(a[:, 7], out["time"], rtol=1e-14)
    _, rows = _csv(os.path.join(ref, "SR-results.csv"))
    want = [float(x) for x in rows[0][1:5]]
    np.testing.assert_allclose(out["tau"] + [out["rate"]], want, rtol=1e-10)


def test_leaf_vectors_vs_reference():
    """pdist_ortho (bit-exact incl. half-box ties and x == L), wiener_khinchin (with NaN
    interpolation) and the Y2m/F2m ufuncs against outputs of the reference's own functions."""
    d = np.load(os.path.join(GOLDEN, "leaf_vectors.npz"))
    for k in range(4):
        x, y, z = d["pd%d_in" % k]
        L, dmax = d["pd%d_par" % k]
        got = orc.pdist_ortho(x, y, z, L, L, L, d["pd%d_idx" % k], dmax)
        for c, nm in enumerate(("dx", "dy", "dz", "dr", "a0", "a1", "prj")):
            np.testing.assert_array_equal(got[c], d["pd%d_%s" % (k, nm)])
    np.testing.assert_allclose(orc.wiener_khinchin(d["wk_in"]), d["wk_out"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(orc.wiener_khinchin(d["wk_nan_in"]), d["wk_nan_out"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(orc.wiener_khinchin(d["wk_real_in"]), d["wk_real_out"], rtol=0, atol=1e-14)
    dx, dy, dz, dr = d["dd_in"]
    y20, y21, y22, r3, f0, f1, f2 = orc.dd_words(dx, dy, dz, dr)
    for got, nm in ((r3, "r-3"), (y20, "Y20"), (y21, "Y21"), (y22, "Y22"), (f0, "F20"), (f1, "F21"), (f2, "F22")):
        np.testing.assert_allclose(got, d["dd_" + nm], rtol=1e-14, atol=1e-16)


# ---------------------------------------------------------------- J(omega) / cutoff / dx
SPECDENS_CASES = 4


def specdens_golden():
    return np.load(os.path.join(GOLDEN, "specdens_vectors.npz"))


def test_specdens_vectors_vs_reference():
    g = specdens_golden()
    for k in range(SPECDENS_CASES):
        t, Y = g["t%d" % k], g["Y%d" % k]
        J = orc.dd_spec_dens(t, Y[0], Y[1], Y[2])
        ref = g["J%d" % k]
        np.testing.assert_allclose(J["omega"], ref[0].real, rtol=1e-15)
        np.testing.assert_allclose(J["sqrt_omega"], ref[2].real, rtol=1e-15)
        assert np.isneginf(J["log_omega"][0]) and np.isneginf(ref[1].real[0])
        np.testing.assert_allclose(J["log_omega"][1:], ref[1].real[1:], rtol=1e-15)
        for c, nm in enumerate(("J0", "J1", "J2")):
            np.testing.assert_allclose(J[nm], ref[3 + c], rtol=0, atol=1e-13 * np.abs(ref[3 + c]).max())
        for tag, kw in (("plain", {}), ("dt", dict(dt=0.002)), ("cut", dict(cutoff=True)),
                        ("cut2", dict(cutoff=True, cutoff_tol=2e-2))):
            j, _ = orc.integrate_columns(t, Y[:3], **kw)
            want = g["dd_%s%d" % (tag, k)]
            np.testing.assert_allclose(2 * j, want[:3], rtol=1e-12, err_msg="dd %s %d" % (tag, k))
            tau_c = 0.5 * (2 * j[1] / Y[1, 0] + 2 * j[2] / Y[2, 0])
            np.testing.assert_allclose(tau_c, want[3], rtol=1e-12)
            q, _ = orc.integrate_columns(t, Y, **kw)
            wq = g["qr_%s%d" % (tag, k)]
            np.testing.assert_allclose(q, wq[:5], rtol=1e-12, err_msg="qr %s %d" % (tag, k))
            np.testing.assert_allclose(q.mean(), wq[5], rtol=1e-12)


def test_cutoff_bound_semantics():
    # the bound of the LAST column is the one applied to every column; short tables fall back to n-1
    g = specdens_golden()
    t, Y = g["t0"], g["Y0"]
    _, b3 = orc.integrate_columns(t, Y[:3], cutoff=True)
    _, b5 = orc.integrate_columns(t, Y, cutoff=True)
    assert 99 <= b3[-1] < len(t) - 1 and 99 <= b5[-1] < len(t) - 1
    assert len(set(b5)) > 1
    assert orc.cutoff_bound(g["Y3"][0]) == len(g["t3"]) - 1
    y = np.ones(300); y[150:] = 0.0
    # trailing 100-mean first reaches <= 1e-3 when the window holds only zeros: i = 249
    assert orc.cutoff_bound(y) == 249
    assert orc.cutoff_bound(y, tol=0.5) == 199


def test_finite_freq_formula_as_written():
    # ddrax.py:249-262 evaluates  b - m*sqrt(omega0)  with m the fitted SLOPE (not its magnitude): for
    # J(omega) = B + S sqrt(omega) the as-written result is B - S sqrt(omega0).  Restated literally.
    w = np.arange(10) * 1.0e5
    J = dict(sqrt_omega=np.sqrt(w), J0=3.0 - 2e-4 * np.sqrt(w) + 0j, J1=2.0 - 1e-4 * np.sqrt(w) + 0j,
             J2=1.0 - 5e-5 * np.sqrt(w) + 0j)
    r = orc.dd_finite_freq(J, larmor_freq=25)
    np.testing.assert_allclose([r["j0"], r["j02"], r["j1"], r["j2"]],
                               [3.0 + 2e-4 * 5e3, 3.0 + 2e-4 * np.sqrt(5e7), 2.0 + 1e-4 * 5e3, 1.0 + 5e-5 * np.sqrt(5e7)],
                               rtol=1e-12)
    assert r["R"] > 0 and r["zf"] > 0


# ---------------------------------------------------------------- nearest-molecule clusters
NN_CASES = [("ion", [24], [0, 2, 4], dict(Na=0.1)), ("lab0", [0], [2], {}), ("lab0_4", [0, 4], [3], dict(Na=0.1)),
            ("symO", "O", [1], dict(Na=0.1))]


def nn_reference(traj, tag, sizes):
    ref = os.path.join(GOLDEN, "nn_water", "ref")
    _, near = _csv(os.path.join(ref, "%s_%s_nearest.csv" % (traj, tag)))
    near = np.array([[int(v) for v in r] for r in near], dtype=np.int64).reshape(-1, 5)
    clusters = {}
    for nn in sizes:
        _, rows = _csv(os.path.join(ref, "%s_%s_nn%d.csv" % (traj, tag, nn)))
        clusters[nn] = [(r[0], float(r[1]), float(r[2]), float(r[3]), int(r[4])) for r in rows]
    return near, clusters


@pytest.mark.parametrize("traj", ["01", "02"])
def test_nearest_neighbors_vs_reference(traj):
    PD = load_params("nn_water").ParseDynamics
    pos, sym, frames = orc.read_xyz(os.path.join(GOLDEN, "nn_water", traj, "traj.xyz"), PD.nat, PD.start_prod,
                                    PD.end_prod, PD.sample_freq, PD.md_print_freq)
    for tag, source, sizes, kw in NN_CASES:
        near, clusters = orc.nearest_neighbors(pos, sym, frames, source, PD.celldm, sizes, dmax=PD.celldm / 2, **kw)
        rnear, rclusters = nn_reference(traj, tag, sizes)
        np.testing.assert_array_equal(near, rnear, err_msg="%s %s nearest" % (traj, tag))
        for nn in sizes:
            assert [c[0] for c in clusters[nn]] == [c[0] for c in rclusters[nn]]
            assert [c[4] for c in clusters[nn]] == [c[4] for c in rclusters[nn]]
            got = np.array([c[1:4] for c in clusters[nn]])
            want = np.array([c[1:4] for c in rclusters[nn]])
            np.testing.assert_array_equal(got, want, err_msg="%s %s nn%d coordinates" % (traj, tag, nn))
    if traj == "02":   # the wrapped trajectory must exercise molecules bonded across periodic images
        _, clusters = orc.nearest_neighbors(pos, sym, frames, [24], PD.celldm, [4], dmax=PD.celldm / 2, Na=0.1)
        xyz = np.array([c[1:4] for c in clusters[4]])
        assert (xyz < 0).any() or (xyz >= PD.celldm).any()
