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


def _qr_case(name, nnuc, nframes, ntraj, seed, analyte="127I", symbol="I", with_traj0=False):
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


def main():
    assert refshim.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    # the reference's own (only) known-answer set
    ref_nb = os.path.join(refshim.REFERENCE, "notebooks")
    dst = os.path.join(GOLD, "reference_iodide")
    os.makedirs(dst, exist_ok=True)
    for f in ("I-ADF-efg-I-acfs.csv", "I-ADF-efg-I-relax.csv"):
        shutil.copy(os.path.join(ref_nb, f), os.path.join(dst, f))

    _qr_case("qr_small", nnuc=3, nframes=64, ntraj=2, seed=1, with_traj0=True)
    _qr_case("qr_odd", nnuc=2, nframes=81, ntraj=1, seed=11)

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

    sr_out = ["{traj}/molax.csv", "{traj}/ang_vel_molax.csv", "{traj}/J_cart.csv",
              "{traj}/Jacfs_all.csv", "{traj}/Jacf.csv", "SR-results.csv"]
    m6 = synth.methane_box(6, seed=3)
    sr = "class SpinRotation:\n    mol_type = %r\n    nmol = %d\n    C_SR = %r\n"
    _traj_case("sr_methane", m6, 36, 2, 35, 1, 2, 0.001, sr % ("methane", 5, [16.495, 16.495, -1.875]), "-s", sr_out)
    _traj_case("sr_water", w8, 30, 1, 29, 2, 1, 0.001, sr % ("water", 8, [36.9, 33.5, 35.5]), "-s", sr_out)


if __name__ == "__main__":
    main()
