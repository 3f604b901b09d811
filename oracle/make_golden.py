"""TEST INFRASTRUCTURE ONLY — regenerates tests/golden/* by running the UNMODIFIED
reference (through oracle/refshim.py) on small seeded synthetic inputs.

    python oracle/make_golden.py            # needs /root/reference; a few minutes

Each case directory holds the exact inputs the reference read (``params.py``, the
``01/*.xyz`` trajectory or the EFG CSV) and, under ``ref/``, the CSVs it wrote.
The GPU box has no /root/reference: tests only read these committed fixtures.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refshim  # noqa: E402
from dynpy_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

PD_TMPL = """class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = {trajs!r}
    nat = {nat}
    start_prod = {start}
    end_prod = {end}
    md_print_freq = {mdp}
    sample_freq = {samp}
    celldm = {celldm!r}
    timestep = {dt!r}
    parse_vel = False
"""


def _traj_case(name, box, nprinted, start, end, samp, mdp, dt, extra_cls, flag, outputs, trajs=("01",)):
    case = os.path.join(GOLD, name)
    shutil.rmtree(case, ignore_errors=True)
    os.makedirs(case)
    for i, tr in enumerate(trajs):
        os.makedirs(os.path.join(case, tr))
        traj = box.frames(i * 1000, i * 1000 + nprinted)
        synth.write_xyz(os.path.join(case, tr, "traj.xyz"), traj, box.symbols)
    body = PD_TMPL.format(trajs=list(trajs), nat=box.nat, start=start, end=end, mdp=mdp, samp=samp,
                          celldm=box.celldm, dt=dt) + extra_cls
    synth.write_params(os.path.join(case, "params.py"), body)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "w")
        shutil.copytree(case, work)
        out = refshim.run_cli(flag, "params.py", cwd=work)
        os.makedirs(os.path.join(case, "ref"))
        with open(os.path.join(case, "ref", "stdout.txt"), "w") as fh:
            fh.write(out)
        for rel in outputs:
            for tr in trajs:
                r = rel.format(traj=tr)
                dst = os.path.join(case, "ref", r.replace("/", "__"))
                shutil.copy(os.path.join(work, r), dst)
    print("golden:", name)


class _WithSpectator:
    """A synthetic box plus one unbonded atom at the emptiest grid point (monatomic species exercise the
    molecule numbering of compute_molecule, universe.py:426-437)."""

    def __init__(self, box, symbol, nprobe):
        import numpy as np
        import torch
        self.box, self.celldm = box, box.celldm
        self.symbols = box.symbols + [symbol]
        self.nat = box.nat + 1
        L = box.celldm
        g = np.stack(np.meshgrid(*(np.linspace(0, L, 16, endpoint=False),) * 3, indexing="ij"), -1).reshape(-1, 3)
        at = box.frames(0, nprobe).permute(2, 0, 1).reshape(-1, 3).numpy()
        d = g[:, None, :] - at[None, :, :]
        d -= L * np.round(d / L)
        self.hole = torch.tensor(g[np.argmax(np.sqrt((d ** 2).sum(-1)).min(1))], dtype=torch.float64)

    def frames(self, f0, f1):
        import torch
        t = torch.arange(f0, f1, dtype=torch.float64)
        extra = self.hole[None, :, None] + 0.05 * torch.sin(0.3 * t)[None, None, :]
        return torch.cat([self.box.frames(f0, f1), extra.expand(1, 3, f1 - f0)], 0)


def _qr_case(name, nnuc, nframes, ntraj, seed, analyte="127I", symbol="I", with_traj0=False, nan_cells=None):
    case = os.path.join(GOLD, name)
    shutil.rmtree(case, ignore_errors=True)
    os.makedirs(case)
    efg = synth.efg_series(nnuc, nframes, seed=seed)
    csv = os.path.join(case, "efg.csv")
    synth.write_efg_csv(csv, efg, symbol=symbol, ntraj=ntraj)
    if with_traj0:  # traj 0 rows must be skipped (qrax.py:45-46); add spectator rows of another symbol too
        with open(csv, "a") as fh:
            fh.write("synthetic,0,1000,0.145,0,%s,1.0,0.0,0.0,0.0,-0.5,0.0,0.0,0.0,-0.5\n" % symbol)
            fh.write("synthetic,1,1000,0.145,77,O,1.0,0.0,0.0,0.0,-0.5,0.0,0.0,0.0,-0.5\n")
    if nan_cells:   # empty cells -> NaN in read_csv; wiener_khinchin interpolates Re and Im of each R_2m separately
        with open(csv) as fh:
            lines = fh.read().split("\n")
        hdr = lines[0].split(",")
        for row, cols in nan_cells:
            t = lines[1 + row].split(",")
            for c in cols:
                t[hdr.index(c)] = ""
            lines[1 + row] = ",".join(t)
        with open(csv, "w") as fh:
            fh.write("\n".join(lines))
    synth.write_params(os.path.join(case, "params.py"),
                       "class Quadrupolar:\n    data_set = './efg.csv'\n    analyte = %r\n" % analyte)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "w")
        shutil.copytree(case, work)
        out = refshim.run_cli("-q", "params.py", cwd=work)
        os.makedirs(os.path.join(case, "ref"))
        with open(os.path.join(case, "ref", "stdout.txt"), "w") as fh:
            fh.write(out)
        for f in ("efg-%s-relax.csv" % symbol, "efg-%s-acfs.csv" % symbol):
            shutil.copy(os.path.join(work, f), os.path.join(case, "ref", f))
    print("golden:", name)


def _leaf_vectors():
    """Outputs of the reference's leaf numerics (numba pdist_ortho, wiener_khinchin, Y2m
    ufuncs, SR_func1 is covered by the SR cases) on seeded inputs -> leaf_vectors.npz."""
    import numpy as np
    refshim.import_reference()
    import distances, qrax, ddrax  # the reference's modules
    rng = np.random.default_rng(5)
    out = {}
    for k, (n, L, dmax) in enumerate([(9, 4.0, 2.4), (60, 7.3, 3.1), (60, 7.3, 7.0), (33, 12.5, 40.0)]):
        x, y, z = rng.uniform(0, L, (3, n))
        x[1], y[1], z[1] = np.mod(x[0] + L / 2, L), y[0], z[0]      # exact half-box tie along x
        x[2], y[2], z[2] = x[0], np.mod(y[0] + L / 2, L), np.mod(z[0] + L / 2, L)
        x[3] = L                                                    # np.mod can return L itself
        idx = np.arange(n) + 1000 * k
        res = distances.pdist_ortho(x, y, z, L, L * 1.0, L * 1.0, idx, dmax)
        out["pd%d_in" % k] = np.stack([x, y, z]); out["pd%d_par" % k] = np.array([L, dmax])
        out["pd%d_idx" % k] = idx
        for c, nm in enumerate(("dx", "dy", "dz", "dr", "a0", "a1", "prj")):
            out["pd%d_%s" % (k, nm)] = res[c]
    z = rng.normal(size=50) + 1j * rng.normal(size=50)
    out["wk_in"] = z; out["wk_out"] = qrax.wiener_khinchin(z)
    zn = z.copy(); zn[[3, 4, 20, 48, 49]] = np.nan
    out["wk_nan_in"] = zn; out["wk_nan_out"] = qrax.wiener_khinchin(zn)
    r = rng.normal(size=31)
    out["wk_real_in"] = r; out["wk_real_out"] = ddrax.wiener_khinchin(r)
    d = rng.normal(size=(3, 40)) * 4
    dr = np.sqrt((d ** 2).sum(0))
    class _T(dict):
        __getitem__ = dict.__getitem__
    import pandas as pd
    df = pd.DataFrame({"dx": d[0], "dy": d[1], "dz": d[2], "dr": dr, "time": np.arange(40.0),
                       "label0": 0, "label1": 1})
    F = ddrax.cart_to_dipolar_from_df(df)
    out["dd_in"] = np.vstack([d, dr])
    for nm in ("$r^{-3}$", "$Y_{2,0}$", "$Y_{2,1}$", "$Y_{2,2}$", "$F_{2,0}$", "$F_{2,1}$", "$F_{2,2}$"):
        out["dd_" + nm.strip("$").replace("{", "").replace("}", "").replace("^", "").replace(",", "").replace("_", "")] = \
            np.asarray(F[nm].values)
    np.savez_compressed(os.path.join(GOLD, "leaf_vectors.npz"), **out)
    print("golden: leaf_vectors.npz", sorted(out)[:6], "...")


def _specdens_vectors():
    """Outputs of the reference's finite-frequency / cutoff / dx integration code (ddrax.spec_dens,
    ddrax.spec_dens_extr_narr, qrax.spectral_dens) on seeded synthetic ACF tables -> specdens_vectors.npz.
    The even-n Simpson values are those of the scipy importable here (see the `scipy` entry)."""
    import numpy as np
    import pandas as pd
    import scipy
    refshim.import_reference()
    import qrax, ddrax
    rng = np.random.default_rng(17)
    out = {"scipy": np.array(scipy.__version__)}
    gcols = ['$G_{2,0}$', '$G_{2,1}$', '$G_{2,2}$']
    fcols = ['$f_{2,-2}$', '$f_{2,-1}$', '$f_{2,0}$', '$f_{2,1}$', '$f_{2,2}$']
    jcols = [r'$\omega$', r'$log(\omega)$', r'$\omega^{\\frac{1}{2}}$', r'$J_{0}(\omega)$', r'$J_{1}(\omega)$', r'$J_{2}(\omega)$']
    # n: even / odd / prime / too short for one window; decay and noise chosen so that the cutoff bound
    # falls inside the table for some columns and at the end for others
    for k, (n, tau, noise) in enumerate([(1000, 0.05, 2e-4), (777, 0.3, 1e-3), (1009, 0.02, 5e-3), (60, 0.01, 1e-4)]):
        t = 0.4 + 0.002 * np.arange(n)            # avg-acf.csv time does not start at 0
        amp = rng.uniform(0.5, 2.0, 5)
        taus = tau * rng.uniform(0.6, 1.6, 5)
        Y = amp[:, None] * np.exp(-(t - t[0])[None, :] / taus[:, None]) * np.cos(7.0 * (t - t[0]))[None, :] \
            + noise * rng.normal(size=(5, n))
        out["t%d" % k] = t; out["Y%d" % k] = Y
        dd = pd.DataFrame({"time": t, gcols[0]: Y[0], gcols[1]: Y[1], gcols[2]: Y[2]})
        qr = pd.DataFrame({"time": t, "symbol": "I", **{c: Y[i] for i, c in enumerate(fcols)}})
        with np.errstate(divide="ignore"):
            J = ddrax.spec_dens(dd)
        out["J%d" % k] = np.stack([J[c].to_numpy().astype(np.complex128) for c in jcols])
        for tag, kw in (("plain", {}), ("dt", dict(dt=0.002)), ("cut", dict(cutoff=True)),
                        ("cut2", dict(cutoff=True, cutoff_tol=2e-2))):
            g = ddrax.spec_dens_extr_narr(dd, **kw)
            out["dd_%s%d" % (tag, k)] = np.array([g['$j_{2,0}$'], g['$j_{2,1}$'], g['$j_{2,2}$'], g[r'$\tau_{c}$']],
                                                  dtype=np.float64)
            h = qrax.spectral_dens(qr, **kw)
            out["qr_%s%d" % (tag, k)] = np.array([h[c] for c in ['$g_{2,-2}$', '$g_{2,-1}$', '$g_{2,0}$', '$g_{2,1}$',
                                                                 '$g_{2,2}$', '$g_{iso}$']], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "specdens_vectors.npz"), **out)
    print("golden: specdens_vectors.npz", len(out), "arrays")


def _neighbors_case(name="nn_water"):
    """neighbors.periodic_nearest_neighbors_by_atom of the reference (the cluster extraction of --inputs,
    neighbors.py:22-118) on 8 waters x 3 frames: trajectory 01 raw (molecules whole), 02 the same frames
    wrapped into the box (molecules split across the boundary, bonded through periodic images).
    plus one Na atom.  Sources: the ion's label (custom Na radius so that it stays unbonded), a water atom
    with default radii, two labels, a symbol.  Outputs: ref/<traj>_<tag>_nearest.csv, _nn<k>.csv."""
    import numpy as np
    import torch
    refshim.import_reference()
    import neighbors, parseMD
    case = os.path.join(GOLD, name)
    shutil.rmtree(case, ignore_errors=True)
    os.makedirs(os.path.join(case, "ref"))
    box = synth.water_box(8, seed=2, drift=0.02)
    raw = box.frames(40, 43)
    # a monatomic solute (the usual analyte of --inputs): isolated atoms exercise the molecule numbering
    L = box.celldm
    g = np.stack(np.meshgrid(*(np.linspace(0, L, 25, endpoint=False),) * 3, indexing="ij"), -1).reshape(-1, 3)
    at = raw.permute(2, 0, 1).reshape(-1, 3).numpy()                       # all atoms of all frames
    d = g[:, None, :] - at[None, :, :]
    d -= L * np.round(d / L)
    hole = g[np.argmax(np.sqrt((d ** 2).sum(-1)).min(1))]                   # the emptiest grid point
    ion = torch.tensor(hole, dtype=torch.float64)[None, :, None] \
        + 0.05 * torch.sin(torch.arange(3, dtype=torch.float64) + 1.0)[None, None, :]
    raw = torch.cat([raw, ion.expand(1, 3, 3)], 0)
    symbols = box.symbols + ["Na"]
    for tr, arr in (("01", raw), ("02", torch.remainder(raw, box.celldm))):
        os.makedirs(os.path.join(case, tr))
        synth.write_xyz(os.path.join(case, tr, "traj.xyz"), arr, symbols)
    body = PD_TMPL.format(trajs=["01", "02"], nat=len(symbols), start=1, end=3, mdp=4, samp=1, celldm=box.celldm,
                          dt=0.002)
    synth.write_params(os.path.join(case, "params.py"), body)
    cwd = os.getcwd()
    os.chdir(case)
    try:
        sys.path.insert(0, case)
        import importlib
        params = importlib.import_module("params")
        PD = params.ParseDynamics
        for tr in PD.trajs:
            u, _ = parseMD.PARSE_MD(PD, PD.traj_dir + tr + "/")
            for tag, source, sizes, kw in (("ion", [24], [0, 2, 4], dict(Na=0.1)), ("lab0", [0], [2], {}),
                                           ("lab0_4", [0, 4], [3], dict(Na=0.1)), ("symO", "O", [1], dict(Na=0.1))):
                dct = neighbors.periodic_nearest_neighbors_by_atom(u, source, PD.celldm, sizes, dmax=PD.celldm / 2, **kw)
                dct["nearest"].to_csv(os.path.join("ref", "%s_%s_nearest.csv" % (tr, tag)), index=False)
                for nn in sizes:
                    dct[nn].atom.to_csv(os.path.join("ref", "%s_%s_nn%d.csv" % (tr, tag, nn)), index=False,
                                        float_format="%.17g")
    finally:
        os.chdir(cwd)
        sys.path.remove(case)
    print("golden:", name)


# acetonitrile through the 'methyl' (top: the five methyl_indeces atoms only / global_momentum: whole molecule)
# and 'ring' axis rules (SRrax.py:405-432, SRparse.py:155-163)
_SR_ACN = ("class SpinRotation:\n    mol_type = %r\n    identifier = 'H(3)C(2)N(1)'\n    nmol = 4\n"
           "    C_SR = [1.0, 1.0, 12.0]\n    methyl_indeces = %r\n    global_momentum = %r\n")
_SR_METHYL_CASES = [("sr_methyl", _SR_ACN % ("methyl", [0, 1, 2, 3, 4], False)),
                    ("sr_methyl_global", _SR_ACN % ("methyl", [0, 1, 2, 3, 4], True)),
                    ("sr_ring", _SR_ACN % ("ring", [0, 2], False))]


def _write_tinker(path, arr, symbols, title):
    """Tinker .arc-style blocks: 'nat title' + box line + 'idx sym x y z type conn...' (Angstrom in, the array is
    bohr); every third frame carries Fortran D exponents, which the reference accepts (d_to_e)."""
    ang = arr.numpy() * 0.529177
    with open(path, "w") as fh:
        for f in range(ang.shape[2]):
            fh.write("%6d  %s\n   30.000000   30.000000   30.000000   90.000000   90.000000   90.000000\n"
                     % (len(symbols), title))
            for a, sym in enumerate(symbols):
                x, y, z = ang[a, :, f]
                if f % 3 == 2:
                    nums = ("%16.8E%16.8E%16.8E" % (x, y, z)).replace("E", "D")
                else:
                    nums = "%14.8f%14.8f%14.8f" % (x, y, z)
                fh.write("%6d  %-3s%s %5d %5d\n" % (a + 1, sym, nums, 1 + a % 3, 2))


def _sr_tinker_case(name="sr_tinker_vel"):
    """Spin-rotation of 6 waters read from a Tinker .arc with EXPLICIT velocities (.vel, parse_vel = True):
    parse_tinker_md (parseMD.py:177-228) and the branch of prep_SR_uni1 that skips the finite difference."""
    case = os.path.join(GOLD, name)
    shutil.rmtree(case, ignore_errors=True)
    os.makedirs(os.path.join(case, "01"))
    box = synth.water_box(6, seed=21)
    nprinted = 26
    pos = box.frames(0, nprinted)
    ext = box.frames(-1, nprinted + 1)
    vel = (ext[:, :, 2:] - ext[:, :, :-2]) / (2 * 0.001)                     # central difference, bohr/ps
    _write_tinker(os.path.join(case, "01", "water.arc"), pos, box.symbols, "water box")
    _write_tinker(os.path.join(case, "01", "water.vel"), vel, box.symbols, "velocities")
    body = PD_TMPL.format(trajs=["01"], nat=box.nat, start=2, end=26, mdp=2, samp=2, celldm=box.celldm, dt=0.001)
    body = body.replace("MD_format = 'xyz'", "MD_format = 'Tinker'").replace("parse_vel = False", "parse_vel = True")
    body += "class SpinRotation:\n    mol_type = 'water'\n    nmol = 6\n    C_SR = [36.9, 33.5, 35.5]\n"
    synth.write_params(os.path.join(case, "params.py"), body)
    outputs = ["{traj}/molax.csv", "{traj}/ang_vel_molax.csv", "{traj}/J_cart.csv", "{traj}/Jacfs_all.csv",
               "{traj}/Jacf.csv", "SR-results.csv"]
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "w")
        shutil.copytree(case, work)
        out = refshim.run_cli("-s", "params.py", cwd=work)
        os.makedirs(os.path.join(case, "ref"))
        with open(os.path.join(case, "ref", "stdout.txt"), "w") as fh:
            fh.write(out)
        for rel in outputs:
            r = rel.format(traj="01")
            shutil.copy(os.path.join(work, r), os.path.join(case, "ref", r.replace("/", "__")))
    print("golden:", name)


def _write_qe(path, arr, lead=None, step=10, dt_ps=0.0001):
    """QE (cp.x) .pos / .vel blocks: one header line 'step time' + nat lines 'x y z' in bohr as given (the reference
    does not convert QE coordinates, parseMD.py:67-153); `lead`: one leading token per line (the .vel reader takes
    columns 1-3, :127)."""
    a = arr.numpy()
    with open(path, "w") as fh:
        for f in range(a.shape[2]):
            fh.write("   %d   %.8f\n" % ((f + 1) * step, (f + 1) * step * dt_ps))
            for k in range(a.shape[0]):
                x, y, z = a[k, :, f]
                fh.write("%s %.10f %.10f %.10f\n" % (lead[k] if lead else "", x, y, z))


def _qe_cases():
    """MD_format = 'QE' (parse_qe_md, parseMD.py:67-153): the trajectory directory is `traj` itself, frame ids come
    from the .pos header lines (10, 20, ... here, NOT printed number x md_print_freq), no unit conversion, symbols
    from the parameter class.  sr_qe: 6 waters, sampled frames, velocities by finite difference; dd_qe: 8 waters.
    (parse_vel = True is not a case: the reference raises at parseMD.py:131 for more than one kept frame.)"""
    sr_out = ["{traj}/molax.csv", "{traj}/ang_vel_molax.csv", "{traj}/J_cart.csv", "{traj}/Jacfs_all.csv",
              "{traj}/Jacf.csv", "SR-results.csv"]
    for name, box, nprinted, start, end, samp, flag, extra, outputs in (
            ("sr_qe", synth.water_box(6, seed=21), 26, 2, 26, 2, "-s",
             "class SpinRotation:\n    mol_type = 'water'\n    nmol = 6\n    C_SR = [36.9, 33.5, 35.5]\n", sr_out),
            ("dd_qe", synth.water_box(8, seed=2, drift=0.02), 40, 1, 40, 1, "-d",
             "class Dipolar:\n    identifier = 'H(2)O(1)'\n    pair_type = 'all'\n    rmax = 9.0\n",
             ["{traj}/avg-acf.csv", "DD-results.csv"])):
        case = os.path.join(GOLD, name)
        shutil.rmtree(case, ignore_errors=True)
        os.makedirs(os.path.join(case, "01"))
        _write_qe(os.path.join(case, "01", "md.pos"), box.frames(0, nprinted))
        body = PD_TMPL.format(trajs=["01"], nat=box.nat, start=start, end=end, mdp=7, samp=samp, celldm=box.celldm,
                              dt=0.001)
        body = body.replace("MD_format = 'xyz'", "MD_format = 'QE'") + "    symbols = %r\n" % (box.symbols,) + extra
        synth.write_params(os.path.join(case, "params.py"), body)
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, "w")
            shutil.copytree(case, work)
            out = refshim.run_cli(flag, "params.py", cwd=work)
            os.makedirs(os.path.join(case, "ref"))
            with open(os.path.join(case, "ref", "stdout.txt"), "w") as fh:
                fh.write(out)
            for rel in outputs:
                r = rel.format(traj="01")
                shutil.copy(os.path.join(work, r), os.path.join(case, "ref", r.replace("/", "__")))
        print("golden:", name)


def _dd_dynamic_cases():
    """6 waters in a small box: in one sampled frame two molecules touch (an H-H contact inside the bond
    cut-off), so the per-frame molecule detection merges them there.  'all' pairs do not depend on molecules;
    'inter' / 'intra' pairs change their status in that frame (DDparse.py:57-60 filters pair ROWS)."""
    w6 = synth.water_box(6, seed=21)
    dd = "class Dipolar:\n    identifier = 'H(2)O(1)'\n    pair_type = %r\n    rmax = %r\n"
    for name, ptype in (("dd_dynamic_all", "all"), ("dd_dynamic_inter", "inter"), ("dd_dynamic_intra", "intra")):
        _traj_case(name, w6, 26, 2, 26, 2, 2, 0.001, dd % (ptype, 12.0), "-d", ["{traj}/avg-acf.csv", "DD-results.csv"])


def _dd_hetero_case():
    """13C-1H dipolar relaxation of the methyl carbon of acetonitrile: analyte_label = 1 (mol-atom index of the
    methyl C), unlike-spin rate formula (ddrax.py:270-271), intramolecular pairs only."""
    dd_het = ("class Dipolar:\n    identifier = 'H(3)C(2)N(1)'\n    pair_type = 'intra'\n    rmax = 8.0\n"
              "    analyte_label = 1\n    analyte_species = '13C'\n")
    _traj_case("dd_hetero_13C", synth.acetonitrile_box(4, seed=4), 40, 1, 40, 1, 2, 0.002, dd_het, "-d",
               ["{traj}/avg-acf.csv", "DD-results.csv"])


# missing EFG values (rows are frame-major, 2 nuclei per frame): a whole tensor, single components (so that
# Re R_22 = (Vxx - Vyy)/2 is missing where only ONE of its terms is), a run of frames, the last frame
_ALLV = ["Vxx", "Vxy", "Vxz", "Vyx", "Vyy", "Vyz", "Vzx", "Vzy", "Vzz"]
_QR_NAN = dict(nnuc=2, nframes=60, ntraj=1, seed=23,
               nan_cells=[(10, _ALLV), (21, ["Vxx"]), (24, ["Vyy", "Vxz", "Vzx"]), (31, ["Vzz"]), (40, _ALLV), (42, _ALLV),
                          (44, _ALLV), (57, ["Vxy", "Vyx"]), (118, _ALLV), (119, ["Vyz", "Vzy"])])


def main():
    assert refshim.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    if "--leaf-only" in sys.argv:
        return _leaf_vectors()
    if "--specdens-only" in sys.argv:
        return _specdens_vectors()
    if "--neighbors-only" in sys.argv:
        return _neighbors_case()
    if "--sr-methyl-only" in sys.argv:
        sr_out = ["{traj}/molax.csv", "{traj}/ang_vel_molax.csv", "{traj}/J_cart.csv",
                  "{traj}/Jacfs_all.csv", "{traj}/Jacf.csv", "SR-results.csv"]
        for name, cls in _SR_METHYL_CASES:
            _traj_case(name, synth.acetonitrile_box(4, seed=4), 22, 1, 22, 1, 1, 0.001, cls, "-s", sr_out)
        return
    if "--sr-tinker-only" in sys.argv:
        return _sr_tinker_case()
    if "--qr-nan-only" in sys.argv:
        return _qr_case("qr_nan", **_QR_NAN)
    if "--dd-dynamic-only" in sys.argv:
        return _dd_dynamic_cases()
    if "--qe-only" in sys.argv:
        return _qe_cases()
    if "--dd-hetero-only" in sys.argv:
        return _dd_hetero_case()
    if "--sr-ar-only" in sys.argv:
        sr_out = ["{traj}/molax.csv", "{traj}/ang_vel_molax.csv", "{traj}/J_cart.csv",
                  "{traj}/Jacfs_all.csv", "{traj}/Jacf.csv", "SR-results.csv"]
        sr = "class SpinRotation:\n    mol_type = %r\n    nmol = %d\n    C_SR = %r\n"
        return _traj_case("sr_methane_ar", _WithSpectator(synth.methane_box(4, seed=13), "Ar", 24), 24, 1, 24, 1, 1,
                          0.001, sr % ("methane", 4, [16.495, 16.495, -1.875]), "-s", sr_out)
    _leaf_vectors()
    _specdens_vectors()
    _neighbors_case()
    # the reference's own (only) known-answer set
    ref_nb = os.path.join(refshim.REFERENCE, "notebooks")
    dst = os.path.join(GOLD, "reference_iodide")
    os.makedirs(dst, exist_ok=True)
    for f in ("I-ADF-efg-I-acfs.csv", "I-ADF-efg-I-relax.csv"):
        shutil.copy(os.path.join(ref_nb, f), os.path.join(dst, f))

    _qr_case("qr_small", nnuc=3, nframes=64, ntraj=2, seed=1, with_traj0=True)
    _qr_case("qr_odd", nnuc=2, nframes=81, ntraj=1, seed=11)
    _qr_case("qr_nan", **_QR_NAN)

    w8 = synth.water_box(8, seed=2, drift=0.02)
    dd = "class Dipolar:\n    identifier = 'H(2)O(1)'\n    pair_type = %r\n    rmax = %r\n"
    dd_out = ["{traj}/avg-acf.csv", "DD-results.csv"]
    _traj_case("dd_gapless_all", w8, 48, 3, 42, 1, 5, 0.002, dd % ("all", 16.0), "-d", dd_out)
    _traj_case("dd_gaps_inter", w8, 48, 3, 42, 1, 5, 0.002, dd % ("inter", 9.0), "-d", dd_out)
    _traj_case("dd_intra_sampled", w8, 48, 2, 46, 2, 1, 0.004, dd % ("intra", 15), "-d", dd_out,
               trajs=("01", "02"))
    dd_het = ("class Dipolar:\n    identifier = 'H(2)O(1)'\n    pair_type = 'all'\n    rmax = 16.0\n"
              "    analyte_label = 1\n    analyte_species = '1H'\n")
    _traj_case("dd_analyte_label", w8, 30, 1, 30, 1, 1, 0.002, dd_het, "-d", dd_out)

    _dd_hetero_case()
    _dd_dynamic_cases()
    _sr_tinker_case()
    _qe_cases()

    sr_out = ["{traj}/molax.csv", "{traj}/ang_vel_molax.csv", "{traj}/J_cart.csv",
              "{traj}/Jacfs_all.csv", "{traj}/Jacf.csv", "SR-results.csv"]
    m6 = synth.methane_box(6, seed=3)
    sr = "class SpinRotation:\n    mol_type = %r\n    nmol = %d\n    C_SR = %r\n"
    _traj_case("sr_methane", m6, 36, 2, 35, 1, 2, 0.001, sr % ("methane", 5, [16.495, 16.495, -1.875]), "-s", sr_out)
    _traj_case("sr_methane_ar", _WithSpectator(synth.methane_box(4, seed=13), "Ar", 24), 24, 1, 24, 1, 1, 0.001,
               sr % ("methane", 4, [16.495, 16.495, -1.875]), "-s", sr_out)
    for name, cls in _SR_METHYL_CASES:
        _traj_case(name, synth.acetonitrile_box(4, seed=4), 22, 1, 22, 1, 1, 0.001, cls, "-s", sr_out)
    _traj_case("sr_water", w8, 30, 1, 29, 2, 1, 0.001, sr % ("water", 8, [36.9, 33.5, 35.5]), "-s", sr_out)


if __name__ == "__main__":
    main()
