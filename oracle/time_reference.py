"""TEST / MEASUREMENT INFRASTRUCTURE ONLY — CPU-baseline provenance.

Times the UNMODIFIED reference (through oracle/refshim.py) in THIS container on the BASELINE.md §3.1 sizes,
then the numpy port (oracle/oracle_np.py, the `cpu_baseline` of bench.py) on exactly the same inputs, checks
that the two agree, and writes units/s, cores and the port-to-reference factor to
profiles/r02_cpu_reference_shim.json.  bench.py copies that record into its `cpu_baseline` object
(`reference_shim_recorded`), because /root/reference does not exist on the GPU box.

    python oracle/time_reference.py [--small]        # needs /root/reference; ~20 min (--small: ~2 min)
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refshim  # noqa: E402
import oracle_np as orc  # noqa: E402
from dynpy_b200 import synth  # noqa: E402
from make_golden import PD_TMPL  # noqa: E402


def _csv_cols(path, cols):
    import csv
    with open(path) as fh:
        rd = list(csv.reader(fh))
    return np.array([[float(r[c]) for c in cols] for r in rd[1:]])


def dd_case(nmol, nframes, work):
    box = synth.water_box(nmol, seed=2, drift=0.02)
    traj = box.frames(0, nframes)
    os.makedirs(os.path.join(work, "01"))
    synth.write_xyz(os.path.join(work, "01", "traj.xyz"), traj, box.symbols)
    rmax = float(box.celldm)          # > half diagonal: gapless, like the bench workloads
    body = PD_TMPL.format(trajs=["01"], nat=box.nat, start=1, end=nframes, mdp=1, samp=1, celldm=box.celldm, dt=0.001)
    body += "class Dipolar:\n    identifier = 'H(2)O(1)'\n    pair_type = 'all'\n    rmax = %r\n" % rmax
    synth.write_params(os.path.join(work, "params.py"), body)
    t0 = time.perf_counter()
    refshim.run_cli("-d", "params.py", cwd=work, timeout=6 * 3600)
    t_ref = time.perf_counter() - t0
    # the port on the inputs the reference read (the xyz text round-trips through 10 decimals)
    pos, sym, frames = orc.read_xyz(os.path.join(work, "01", "traj.xyz"), box.nat, 1, nframes, 1, 1)
    t0 = time.perf_counter()
    out = orc.dd_pipeline(pos, sym, frames, box.celldm, 0.001, rmax=rmax)
    t_port = time.perf_counter() - t0
    # the bench's cpu_baseline calls dd_pairs_gapless (same arithmetic without the pair table): time it too
    w = np.mod(pos.transpose(1, 2, 0), box.celldm)
    hidx = [k for k, s in enumerate(sym) if s == "H"]
    wh = np.ascontiguousarray(w[hidx])
    i, j = np.triu_indices(len(hidx), k=1)
    t0 = time.perf_counter()
    orc.dd_pairs_gapless(wh, list(zip(i.tolist(), j.tolist())), box.celldm)
    t_port2 = time.perf_counter() - t0
    ref = _csv_cols(os.path.join(work, "01", "avg-acf.csv"), range(2, 9)).T
    err = max(float(np.max(np.abs(out["G"][c] - ref[c])) / np.max(np.abs(ref[4:7]))) for c in range(7))
    units = len(i) * nframes
    return {"case": "DDparse.DD_module_main, %d H2O (%d H-H pairs) x %d frames, pair_type all, gapless rmax" % (nmol, len(i), nframes),
            "unit": "pair-frames/s", "units": units, "reference_shim_s": t_ref, "reference_shim_units_per_s": units / t_ref,
            "reference_cores": 1, "port_pipeline_s": t_port, "port_pipeline_units_per_s": units / t_port,
            "port_pairs_gapless_s": t_port2, "port_units_per_s": units / t_port2, "port_cores": 1,
            "port_over_reference": t_ref / t_port2, "max_normwise_diff_port_vs_reference": err}


def qr_case(nnuc, nframes, work):
    efg = synth.efg_series(nnuc, nframes, seed=1)
    synth.write_efg_csv(os.path.join(work, "efg.csv"), efg, symbol="I", ntraj=1)
    synth.write_params(os.path.join(work, "params.py"),
                       "class Quadrupolar:\n    data_set = './efg.csv'\n    analyte = '127I'\n")
    t0 = time.perf_counter()
    refshim.run_cli("-q", "params.py", cwd=work, timeout=6 * 3600)
    t_ref = time.perf_counter() - t0
    rows = orc.read_efg_csv(os.path.join(work, "efg.csv"))
    t0 = time.perf_counter()
    res, acfs = orc.qr_pipeline(rows, "127I")
    t_port = time.perf_counter() - t0
    ref = _csv_cols(os.path.join(work, "efg-I-acfs.csv"), range(2, 7)).T
    got = acfs[1][2]
    err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    units = nnuc * nframes
    return {"case": "qrax.QR_module_main, %d nuclei x %d frames (reference: whole CLI process incl. imports and read_csv; port: after the CSV is parsed)" % (nnuc, nframes),
            "unit": "nucleus-frames/s", "units": units, "reference_shim_s": t_ref, "reference_shim_units_per_s": units / t_ref,
            "reference_cores": 1, "port_s": t_port, "port_units_per_s": units / t_port, "port_cores": 1,
            "port_over_reference": t_ref / t_port, "max_normwise_diff_port_vs_reference": err}


def sr_case(nmol, nframes, work):
    box = synth.methane_box(nmol, seed=3)
    traj = box.frames(0, nframes + 1)
    os.makedirs(os.path.join(work, "01"))
    synth.write_xyz(os.path.join(work, "01", "traj.xyz"), traj, box.symbols)
    body = PD_TMPL.format(trajs=["01"], nat=box.nat, start=1, end=nframes + 1, mdp=1, samp=1, celldm=box.celldm, dt=0.001)
    body += "class SpinRotation:\n    mol_type = 'methane'\n    nmol = %d\n    C_SR = [16.495, 16.495, -1.875]\n" % nmol
    synth.write_params(os.path.join(work, "params.py"), body)
    t0 = time.perf_counter()
    refshim.run_cli("-s", "params.py", cwd=work, timeout=6 * 3600)
    t_ref = time.perf_counter() - t0
    pos, sym, frames = orc.read_xyz(os.path.join(work, "01", "traj.xyz"), box.nat, 1, nframes + 1, 1, 1)
    t0 = time.perf_counter()
    out = orc.sr_pipeline(pos, sym, frames, box.celldm, 0.001, "methane", nmol, [16.495, 16.495, -1.875])
    t_port = time.perf_counter() - t0
    ref = _csv_cols(os.path.join(work, "SR-results.csv"), range(1, 5))[0]
    got = np.array([out["tau"][0], out["tau"][1], out["tau"][2], out["rate"]]) if "tau" in out else None
    err = float(np.max(np.abs(got - ref) / np.abs(ref))) if got is not None else None
    units = nmol * nframes
    return {"case": "SRparse.SR_module_main, %d CH4 x %d frames (Pool(cpu_count()) in the reference)" % (nmol, nframes),
            "unit": "molecule-frames/s", "units": units, "reference_shim_s": t_ref, "reference_shim_units_per_s": units / t_ref,
            "reference_cores": os.cpu_count(), "port_s": t_port, "port_units_per_s": units / t_port, "port_cores": 1,
            "port_over_reference": t_ref / t_port, "max_rel_diff_rates_port_vs_reference": err}


def main():
    assert refshim.available(), "needs /root/reference"
    small = "--small" in sys.argv
    sizes = dict(dd=(8, 200), qr=(2, 4096), sr=(8, 100)) if small else dict(dd=(32, 2000), qr=(8, 65536), sr=(64, 1000))
    import platform
    rec = {"what": "reference through oracle/refshim.py vs the numpy port, same inputs, build container (no GPU)",
           "host": {"cpus": os.cpu_count(), "machine": platform.machine()}, "sizes": "small" if small else "BASELINE.md 3.1",
           "cases": {}}
    for name, fn in (("qr", qr_case), ("sr", sr_case), ("dd", dd_case)):
        with tempfile.TemporaryDirectory() as tmp:
            rec["cases"][name] = fn(*sizes[name], tmp)
        print(name, json.dumps(rec["cases"][name]), flush=True)
        out = os.path.join(ROOT, "profiles", "r02_cpu_reference_shim%s.json" % ("_small" if small else ""))
        with open(out, "w") as fh:
            json.dump(rec, fh, indent=1)


if __name__ == "__main__":
    main()
