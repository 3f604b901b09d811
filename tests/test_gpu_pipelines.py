"""End-to-end parity on the GPU: the drop-in drivers (dynpy_b200.qrax / DDparse / SRparse
and the ``python -m dynpy`` CLI) against the CSVs the UNMODIFIED reference wrote for the
same inputs (tests/golden/*/ref), plus kernel-level DD / SR checks against the oracle.

Tolerances: ACF columns normwise <= 1e-9, rates/taus <= 1e-8 relative, integer columns and
headers exact."""
import csv
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_params
from oracle import oracle_np as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dynpy_b200 import ops as o
    return o


def _csv(path):
    with open(path) as fh:
        rd = list(csv.reader(fh))
    return rd[0], rd[1:]


def _num(rows, cols):
    return np.array([[float(r[c]) if r[c] != "" else np.nan for c in cols] for r in rows])


def nerr(got, ref):
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)


def _workdir(tmp_path, case):
    w = tmp_path / case
    shutil.copytree(os.path.join(GOLDEN, case), w, ignore=shutil.ignore_patterns("ref"))
    return str(w)


def _run_cli(flag, work, stdin="y\n"):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT
    r = subprocess.run([sys.executable, "-m", "dynpy", flag, "params.py"], cwd=work, input=stdin, text=True,
                       capture_output=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


# ------------------------------------------------------------------ QR
@pytest.mark.parametrize("case", ["qr_small", "qr_odd", "qr_nan"])
def test_cli_qrelax_vs_reference_csv(tmp_path, case, ops):
    work = _workdir(tmp_path, case)
    out = _run_cli("-q", work)
    assert "Results written to" in out
    for name, ncol in (("efg-I-relax.csv", 7), ("efg-I-acfs.csv", 8)):
        h1, r1 = _csv(os.path.join(work, name))
        h0, r0 = _csv(os.path.join(GOLDEN, case, "ref", name))
        assert h1 == h0
        assert len(r1) == len(r0)
        if name.endswith("relax.csv"):
            assert [r[:2] for r in r1] == [r[:2] for r in r0]
            a1, a0 = _num(r1, range(2, 7)), _num(r0, range(2, 7))
            ok = ~np.isnan(a0)
            assert (np.isnan(a1) == np.isnan(a0)).all()
            np.testing.assert_allclose(a1[ok], a0[ok], rtol=1e-8, atol=1e-13)
        else:
            a1, a0 = _num(r1, range(0, 7)), _num(r0, range(0, 7))
            np.testing.assert_allclose(a1[:, :2], a0[:, :2], rtol=1e-12, atol=1e-15)
            for c in range(2, 7):
                assert nerr(a1[:, c], a0[:, c]) < 1e-9
            assert [r[7] for r in r1] == [r[7] for r in r0]


def test_qr_golden_iodide_rates_on_gpu(ops):
    """Simpson + rate algebra on the device against the reference's shipped known answer."""
    from dynpy_b200 import qrax
    _, rows = _csv(os.path.join(GOLDEN, "reference_iodide", "I-ADF-efg-I-acfs.csv"))
    a = _num(rows, range(7))
    t = torch.tensor(a[:, 1]).cuda()
    f = torch.tensor(np.ascontiguousarray(a[:, 2:7].T)).cuda()
    g, giso = qrax.spectral_dens(t, f, "avg")
    t1, t2, tiso = qrax.relaxation(g, giso, "127I")
    v0 = float(a[0, 2:7].sum())
    _, rr = _csv(os.path.join(GOLDEN, "reference_iodide", "I-ADF-efg-I-relax.csv"))
    want = [float(x) for x in rr[0][2:7]]
    np.testing.assert_allclose([t1, t2, tiso, giso * 5 / v0, v0], want, rtol=1e-10)


# ------------------------------------------------------------------ DD
@pytest.mark.parametrize("case,trajs", [("dd_gapless_all", ["01"]), ("dd_gaps_inter", ["01"]),
                                         ("dd_intra_sampled", ["01", "02"]), ("dd_analyte_label", ["01"]),
                                         ("dd_hetero_13C", ["01"]), ("dd_dynamic_all", ["01"]),
                                         ("dd_dynamic_inter", ["01"]), ("dd_dynamic_intra", ["01"]),
                                         ("dd_qe", ["01"])])
def test_cli_ddrelax_vs_reference_csv(tmp_path, case, trajs, ops):
    work = _workdir(tmp_path, case)
    out = _run_cli("-d", work)
    assert "1/T1/nH" in out and "Continue? (Y/n)" in out
    for traj in trajs:
        h1, r1 = _csv(os.path.join(work, traj, "avg-acf.csv"))
        h0, r0 = _csv(os.path.join(GOLDEN, case, "ref", "%s__avg-acf.csv" % traj))
        assert h1 == h0
        assert len(r1) == len(r0)
        a1, a0 = _num(r1, range(9)), _num(r0, range(9))
        np.testing.assert_array_equal(a1[:, 0], a0[:, 0])
        np.testing.assert_allclose(a1[:, 1], a0[:, 1], rtol=1e-13)
        for c in range(2, 9):
            # G_r of rigid intramolecular pairs is rounding noise (r^-3 - mean = 0): judge it on the
            # scale of the r^-3-weighted columns it belongs to
            scale = np.max(np.abs(a0[:, 6:9])) if c == 5 else np.max(np.abs(a0[:, c]))
            err = np.max(np.abs(a1[:, c] - a0[:, c])) / scale
            assert err < 1e-9, (traj, c, err)
    h1, r1 = _csv(os.path.join(work, "DD-results.csv"))
    h0, r0 = _csv(os.path.join(GOLDEN, case, "ref", "DD-results.csv"))
    assert h1 == h0 and [r[0] for r in r1] == [r[0] for r in r0]
    np.testing.assert_allclose(_num(r1, range(1, 4)), _num(r0, range(1, 4)), rtol=1e-8)


def test_dd_kernels_vs_oracle_bigger(ops):
    """27 waters x 300 frames, both the gapless fast path and the ragged path, against the oracle."""
    from dynpy_b200 import DDparse, ddrax, synth, topology
    box = synth.water_box(27, seed=12, drift=0.01, trans_amp=0.15)
    F = 300
    traj = box.frames(0, F)
    pos = traj.numpy().transpose(2, 0, 1)
    frames = np.arange(1, F + 1) * 2
    cell = (box.celldm,) * 3
    wrapped = ops.wrap(traj.cuda().contiguous(), box.celldm)
    np.testing.assert_array_equal(wrapped.cpu().numpy(), np.mod(traj.numpy(), box.celldm))
    for rmax, ptype in ((30.0, "all"), (9.5, "inter")):
        ref = orc.dd_pipeline(pos, box.symbols, frames, box.celldm, 0.001, pair_type=ptype, rmax=rmax)
        topo, _ = topology.build_topology(wrapped, box.symbols, cell, rmax)
        pi, pj = DDparse.select_pairs(box.symbols, topo, 3, ptype, None)
        G, hit, nspins, info = ddrax.pair_acf_sum(wrapped, torch.from_numpy(pi).cuda(), torch.from_numpy(pj).cuda(),
                                                  cell, rmax)
        assert nspins == ref["nspins"]
        assert int(hit.sum()) == len(ref["time"])
        Gh = (G[:, hit] * 2.0 / nspins).cpu().numpy()
        for c in range(7):
            assert nerr(Gh[c], ref["G"][c]) < 1e-10, (rmax, c, nerr(Gh[c], ref["G"][c]))
        if rmax < 20:
            assert info["partial"] > 0 and info["full"] > 0


@pytest.mark.parametrize("F", [64, 62, 8200])   # 32-byte loads (F % 4 == 0; 8200: three frame chunks, atomic counts) and the scalar form
def test_dd_pair_count_bit_exact_presence(ops, F):
    """Presence (r2 < rmax^2) must agree with pdist_ortho's decision for every pair-frame."""
    from dynpy_b200 import synth
    box = synth.water_box(8, seed=5)
    traj = box.frames(0, F)
    L = box.celldm
    w = np.mod(traj.numpy(), L)
    A = box.nat
    i, j = np.triu_indices(A, k=1)
    rmax = 0.45 * L
    cnt = np.zeros(len(i), dtype=np.int64)
    for f in range(F):
        res = orc.pdist_ortho(w[:, 0, f], w[:, 1, f], w[:, 2, f], L, L, L, np.arange(A), rmax)
        present = set(zip(res[4].tolist(), res[5].tolist()))
        cnt += np.array([(a, b) in present for a, b in zip(i.tolist(), j.tolist())])
    hit = torch.zeros(F, dtype=torch.int32, device="cuda")
    got = ops.dd_pair_count(torch.tensor(w).cuda().contiguous(), torch.tensor(i, dtype=torch.int32).cuda(),
                            torch.tensor(j, dtype=torch.int32).cuda(), (L, L, L), rmax, frame_hit=hit)
    np.testing.assert_array_equal(got.cpu().numpy(), cnt)
    assert int(hit.sum()) == F


def test_topology_change_is_detected(ops):
    """Molecules that come within bonding distance in some frame change the reference's per-frame
    molecule table; the static-topology path must refuse instead of silently mis-indexing."""
    from dynpy_b200 import synth, topology
    box = synth.water_box(27, seed=12, drift=0.01)  # default amplitudes: two waters touch in one frame
    traj = box.frames(0, 300)
    two, mol = orc.atom_two_table(traj.numpy().transpose(2, 0, 1), box.symbols, box.celldm, 9.5)
    assert min(len(set(mol[f])) for f in range(300)) < 27
    wrapped = ops.wrap(traj.cuda().contiguous(), box.celldm)
    with pytest.raises(RuntimeError, match="topology changes"):
        topology.build_topology(wrapped, box.symbols, (box.celldm,) * 3, 9.5)
    # dynamic=True: the device scan lists the deviating pair-frames and the host re-indexes those frames —
    # the molecule number of every atom in every frame must equal the per-frame detection of the oracle
    ft, _ = topology.build_topology(wrapped, box.symbols, (box.celldm,) * 3, 9.5, dynamic=True)
    assert ft.frames and 0 not in ft.frames
    np.testing.assert_array_equal(ft.molecule_ids(np.arange(box.nat)), mol)


# ------------------------------------------------------------------ SR
@pytest.mark.parametrize("case", ["sr_methane", "sr_water", "sr_methane_ar", "sr_methyl", "sr_methyl_global", "sr_ring",
                                  "sr_tinker_vel", "sr_qe"])
def test_cli_srrelax_vs_reference_csv(tmp_path, case, ops):
    work = _workdir(tmp_path, case)
    out = _run_cli("-s", work)
    assert "1/T1     =" in out
    ref = os.path.join(GOLDEN, case, "ref")
    for name, ncols, nint in (("molax", 9, 3), ("ang_vel_molax", 3, 3), ("J_cart", 3, 3), ("Jacfs_all", 3, 3)):
        h1, r1 = _csv(os.path.join(work, "01", name + ".csv"))
        h0, r0 = _csv(os.path.join(ref, "01__%s.csv" % name))
        assert h1 == h0, name
        assert len(r1) == len(r0), name
        a1, a0 = _num(r1, range(1, 1 + ncols + nint)), _num(r0, range(1, 1 + ncols + nint))
        np.testing.assert_array_equal(a1[:, ncols:], a0[:, ncols:], err_msg=name)
        assert [r[0] for r in r1] == [r[0] for r in r0]
        for c in range(ncols):
            assert nerr(a1[:, c], a0[:, c]) < 1e-9, (name, c, nerr(a1[:, c], a0[:, c]))
    h1, r1 = _csv(os.path.join(work, "01", "Jacf.csv"))
    h0, r0 = _csv(os.path.join(ref, "01__Jacf.csv"))
    assert h1 == h0 and len(r1) == len(r0)
    a1, a0 = _num(r1, range(8)), _num(r0, range(8))
    np.testing.assert_array_equal(a1[:, 0], a0[:, 0])
    for c in (1, 2, 3):
        assert nerr(a1[:, c], a0[:, c]) < 1e-9
    np.testing.assert_allclose(a1[:, 4:], a0[:, 4:], rtol=1e-12)
    h1, r1 = _csv(os.path.join(work, "SR-results.csv"))
    h0, r0 = _csv(os.path.join(ref, "SR-results.csv"))
    assert h1 == h0
    np.testing.assert_allclose(_num(r1, range(1, 5)), _num(r0, range(1, 5)), rtol=1e-8)


def test_sr_kernel_vs_oracle_bigger(ops):
    from dynpy_b200 import SRparse, synth

    class PD:
        celldm = synth.methane_celldm(20)
        timestep = 0.002

    class SR:
        mol_type = "methane"
        nmol = 17
        C_SR = [16.495, 16.495, -1.875]

    box = synth.methane_box(20, seed=8)
    F = 120
    traj = box.frames(0, F)
    frames = np.arange(10, 10 + F) * 3
    from dynpy_b200.parseMD import Trajectory
    tr = Trajectory(pos=traj.cuda().contiguous(), symbols=box.symbols, frames=frames)
    out = SRparse.sr_trajectory(tr, PD, SR, "cartwright")
    ref = orc.sr_pipeline(traj.numpy().transpose(2, 0, 1), box.symbols, frames, PD.celldm, PD.timestep, "methane", 17,
                          SR.C_SR)
    nmol = len(out["keep"])
    J = out["J"].permute(2, 0, 1).reshape(-1, 3).cpu().numpy()
    assert J.shape == ref["table"]["J"].shape
    assert nerr(J, ref["table"]["J"]) < 1e-11
    for c in range(3):
        assert nerr(out["acf_mean"][c].cpu().numpy(), ref["acf_mean"][c]) < 1e-11
    np.testing.assert_allclose(list(out["tau"]) + [out["rate"]], ref["tau"] + [ref["rate"]], rtol=1e-9)
    assert nmol == 17


# ------------------------------------------------------------------ CLI contract
def test_cli_missing_key_and_bad_flag(tmp_path):
    work = tmp_path / "w"
    work.mkdir()
    (work / "params.py").write_text("class Quadrupolar:\n    data_set = 'x.csv'\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "dynpy", "-q", "params.py"], cwd=work, text=True, capture_output=True, env=env)
    assert r.returncode == 2 and "Missing required input variable analyte in class Quadrupolar" in r.stdout
    r = subprocess.run([sys.executable, "-m", "dynpy", "--nonsense"], cwd=work, text=True, capture_output=True, env=env)
    assert r.returncode == 2 and "USAGE:" in r.stdout
    r = subprocess.run([sys.executable, "-m", "dynpy"], cwd=work, text=True, capture_output=True, env=env)
    assert r.returncode == 2 and "USAGE:" in r.stdout


# ------------------------------------------------------------------ nearest-molecule clusters
NN_CASES = [("ion", [24], [0, 2, 4], dict(Na=0.1)), ("lab0", [0], [2], {}), ("lab0_4", [0, 4], [3], dict(Na=0.1)),
            ("symO", "O", [1], dict(Na=0.1))]


@pytest.mark.parametrize("traj", ["01", "02"])
def test_neighbors_vs_reference(ops, traj):
    """dynpy_b200.neighbors.periodic_nearest_neighbors_by_atom against the tables the reference wrote:
    every column of `nearest` (frame, idx, two, atom, molecule) and the cluster atoms (symbol, frame exact,
    coordinates bit-exact) — raw (01) and wrapped (02: molecules bonded across periodic images) frames."""
    from dynpy_b200 import neighbors, parseMD
    PD = load_params("nn_water").ParseDynamics
    PD.traj_dir = os.path.join(GOLDEN, "nn_water") + os.sep
    tr = parseMD.PARSE_MD(PD, traj)
    ref = os.path.join(GOLDEN, "nn_water", "ref")
    for tag, source, sizes, kw in NN_CASES:
        dct = neighbors.periodic_nearest_neighbors_by_atom(tr, source, PD.celldm, sizes, dmax=PD.celldm / 2, **kw)
        hdr, rows = _csv(os.path.join(ref, "%s_%s_nearest.csv" % (traj, tag)))
        assert list(dct["nearest"].columns) == hdr
        want = np.array([[int(v) for v in r] for r in rows], dtype=np.int64).reshape(-1, 5)
        np.testing.assert_array_equal(dct["nearest"].to_numpy(dtype=np.int64), want, err_msg="%s %s" % (traj, tag))
        for nn in sizes:
            hdr, rows = _csv(os.path.join(ref, "%s_%s_nn%d.csv" % (traj, tag, nn)))
            atom = dct[nn].atom
            assert list(atom.columns) == hdr
            assert atom["symbol"].tolist() == [r[0] for r in rows]
            assert atom["frame"].tolist() == [int(r[4]) for r in rows]
            np.testing.assert_array_equal(atom[["x", "y", "z"]].to_numpy(), _num(rows, [1, 2, 3]))


def test_neighbors_larger_box_vs_oracle(ops):
    """32 waters + 1 ion (33*27 = 2673 atoms in the image block, 3.6e6 pairs -> the counted pdist path):
    product vs the numpy restatement, wrapped coordinates."""
    from dynpy_b200 import neighbors, parseMD, synth
    box = synth.water_box(32, seed=9, drift=0.05)
    raw = box.frames(100, 102)
    L = box.celldm
    ion = torch.tensor([[0.013 * L, 0.5 * L, 0.48 * L]], dtype=torch.float64)[:, :, None].expand(1, 3, 2)
    raw = torch.remainder(torch.cat([raw, ion], 0), L)
    symbols = box.symbols + ["Na"]
    tr = parseMD.Trajectory(pos=raw.cuda().contiguous(), symbols=symbols, frames=np.array([5, 10], dtype=np.int64))
    dct = neighbors.periodic_nearest_neighbors_by_atom(tr, [96], L, [6], dmax=L / 2, Na=0.1)
    near, clusters = orc.nearest_neighbors(raw.numpy().transpose(2, 0, 1), symbols, [5, 10], [96], L, [6], dmax=L / 2,
                                           Na=0.1)
    np.testing.assert_array_equal(dct["nearest"].to_numpy(dtype=np.int64), near)
    atom = dct[6].atom
    assert atom["symbol"].tolist() == [c[0] for c in clusters[6]]
    np.testing.assert_array_equal(atom[["x", "y", "z"]].to_numpy(), np.array([c[1:4] for c in clusters[6]]))
    assert len(atom) >= 2 * (6 * 3 + 1)


# ------------------------------------------------------------------ sharded runs (2 ranks on one GPU, gloo)
@pytest.mark.parametrize("flag,case,files", [
    ("-d", "dd_gaps_inter", ["01/avg-acf.csv", "DD-results.csv"]),
    ("-d", "dd_dynamic_intra", ["01/avg-acf.csv", "DD-results.csv"]),
    ("-s", "sr_methane", ["01/Jacf.csv", "01/J_cart.csv", "SR-results.csv"]),
    ("-s", "sr_tinker_vel", ["01/Jacf.csv", "01/J_cart.csv", "01/molax.csv", "SR-results.csv"]),
    ("-q", "qr_small", ["efg-I-acfs.csv", "efg-I-relax.csv"])])
def test_cli_two_ranks_equal_one_rank(tmp_path, flag, case, files, ops):
    """The same CLI under torchrun with two ranks (pairs / molecules / nuclei sharded, one all-reduce; gloo so
    that both ranks can share the single GPU of the test box) must write what the single-process run writes."""
    one = _workdir(tmp_path / "one", case)
    _run_cli(flag, one)
    two = _workdir(tmp_path / "two", case)
    env = dict(os.environ, PYTHONPATH=ROOT, DYNPY_DIST_BACKEND="gloo", OMP_NUM_THREADS="1")
    port = 29000 + (os.getpid() * 7 + len(case)) % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "dynpy", flag, "params.py"],
                       cwd=two, input="y\n", text=True, capture_output=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for f in files:
        h1, r1 = _csv(os.path.join(one, f))
        h2, r2 = _csv(os.path.join(two, f))
        assert h1 == h2 and len(r1) == len(r2), f
        for a, b in zip(r1, r2):
            for x, y in zip(a, b):
                try:
                    fx, fy = float(x), float(y)
                except ValueError:
                    assert x == y
                    continue
                if np.isnan(fx) and np.isnan(fy):
                    continue
                assert abs(fx - fy) <= 1e-9 * max(abs(fx), 1e-30) + 1e-13, (f, x, y)  # the `err` row of one trajectory is rounding noise
