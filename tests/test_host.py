"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls), the CLI contract, trajectory parsing, topology/pair indexing
against the oracle, and the multi-rank plumbing (world_size 2, gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_params
from oracle import oracle_np as orc


# ------------------------------------------------------------------ C-ABI surface
def _header_functions():
    src = open(os.path.join(ROOT, "include", "dynpy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dynpy_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dynpy_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dynpy_version() >= 100
    # pure host helpers may be called without a GPU
    assert lib.dynpy_fft_length(316) == 1024
    assert lib.dynpy_fft_length(100000) == 204800  # 50 x 4096 covers 65536 < n <= 102400
    assert lib.dynpy_fft_length(65536) == 131072
    assert lib.dynpy_fft_length(102401) == 262144
    assert lib.dynpy_fft_length(36000) == 20 * 4096 and lib.dynpy_fft_length(45000) == 24 * 4096
    assert lib.dynpy_fft_length(70000) == 40 * 4096 and lib.dynpy_fft_length(32768) == 65536
    assert lib.dynpy_fft_length(1) == 32
    assert lib.dynpy_fft_workspace_bytes(4096, 10) == 0
    assert lib.dynpy_fft_workspace_bytes(8192, 3) == 3 * 8192 * 16


def test_no_cpu_fallback():
    from dynpy_b200 import _lib, ops
    with pytest.raises(_lib.DynpyError, match="CUDA tensor"):
        ops.wk_acf_sum(torch.zeros((2, 16), dtype=torch.float64))
    with pytest.raises(_lib.DynpyError, match="CUDA tensor"):
        ops.wrap(torch.zeros(4, dtype=torch.float64), 3.0)
    if not torch.cuda.is_available():  # a compute entry point without a device must fail loudly
        lib = _lib.load()
        assert lib.dynpy_device_sm_count() < 0
        assert b"CUDA" in lib.dynpy_last_error() or b"device" in lib.dynpy_last_error()


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dynpy_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn
            assert "oracle_np" not in txt and "refshim" not in txt, fn
    assert "oracle" not in open(os.path.join(ROOT, "dynpy.py")).read()


# ------------------------------------------------------------------ CLI contract (dynpy.py:42-124)
def _cli(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-m", "dynpy"] + args, cwd=cwd, text=True, capture_output=True, env=env,
                          timeout=300)


def test_cli_contract(tmp_path):
    (tmp_path / "params.py").write_text("class Quadrupolar:\n    data_set = 'x.csv'\n")
    r = _cli(["-q", "params.py"], tmp_path)
    assert r.returncode == 2 and "Error: Missing required input variable analyte in class Quadrupolar." in r.stdout
    r = _cli(["--nonsense"], tmp_path)
    assert r.returncode == 2 and r.stdout.startswith("USAGE:")
    r = _cli([], tmp_path)
    assert r.returncode == 2 and r.stdout.startswith("USAGE:")
    r = _cli(["-h"], tmp_path)
    assert r.returncode == 0 and "--DDrelax <params.py>" in r.stdout
    r = _cli(["-d", "nofile.py"], tmp_path)
    assert r.returncode == 2 and "Input parameter file must have .py file extension" in r.stdout
    (tmp_path / "dd.py").write_text("class ParseDynamics:\n    MD_format='xyz'\nclass Dipolar:\n    identifier='H(2)O(1)'\n")
    r = _cli(["--DDrax", "dd.py"], tmp_path)  # BASELINE.json spelling is an alias
    assert r.returncode == 2 and "Missing required input variable traj_dir in class ParseDynamics" in r.stdout


# ------------------------------------------------------------------ trajectory ingestion
@pytest.mark.parametrize("case", ["dd_gapless_all", "dd_intra_sampled", "sr_methane"])
def test_parse_xyz_matches_reference_semantics(case):
    from dynpy_b200 import parseMD
    P = load_params(case)
    PD = P.ParseDynamics
    PD.traj_dir = os.path.join(GOLDEN, case) + "/"
    tr = parseMD.PARSE_MD(PD, "01", device=torch.device("cpu"))
    pos, sym, frames = orc.read_xyz(os.path.join(GOLDEN, case, "01", "traj.xyz"), PD.nat, PD.start_prod, PD.end_prod,
                                    PD.sample_freq, PD.md_print_freq)
    assert tr.symbols == sym
    np.testing.assert_array_equal(tr.frames, frames)
    np.testing.assert_array_equal(tr.pos.numpy(), pos.transpose(1, 2, 0))


def test_parse_tinker_arc(tmp_path):
    from dynpy_b200 import parseMD, synth
    box = synth.water_box(4, seed=1)
    traj = box.frames(0, 5).numpy() / synth.BOHR_PER_ANG
    d = tmp_path / "01"
    d.mkdir()
    with open(d / "water.arc", "w") as fh:
        for f in range(5):
            fh.write("%6d  water box\n   20.0 20.0 20.0 90.0 90.0 90.0\n" % box.nat)
            for a in range(box.nat):
                s = ("OW", "HW", "H2")[a % 3]
                x, y, z = traj[a, :, f]
                fh.write(("%6d  %-3s%14.8f%14.8f%14.8f %5d %5d\n" % (a + 1, s, x, y, z, 1, 2)).replace("E", "D"))

    class PD:
        MD_format = "Tinker"; traj_dir = str(tmp_path) + "/"; trajs = ["01"]; nat = box.nat; start_prod = 2
        end_prod = 4; md_print_freq = 10; sample_freq = 2; celldm = 20.0; timestep = 0.001
    tr = parseMD.PARSE_MD(PD, "01", device=torch.device("cpu"))
    assert tr.symbols[:3] == ["O", "H", "H"]
    np.testing.assert_array_equal(tr.frames, [20, 40])
    np.testing.assert_allclose(tr.pos.numpy()[:, :, 0], np.round(traj[:, :, 1], 8) / 0.529177, rtol=0, atol=1e-12)


def test_native_block_reader_matches_python_reader(tmp_path):
    """csrc/ingest.cpp against the pure-Python restatement: bit-identical doubles for plain, signed,
    exponent, Fortran-D and tab/CRLF formatted numbers, frame sub-sampling, no trailing newline."""
    from dynpy_b200 import _lib, parseMD
    rng = np.random.default_rng(5)
    nat, nfr = 7, 23
    path = tmp_path / "t.arc"
    fmts = ["%14.8f", "%+.10e", "%.17g", "%12.6f"]
    with open(path, "w", newline="") as fh:
        for f in range(nfr):
            fh.write("%d title %d\n box line\n" % (nat, f))
            for a in range(nat):
                v = rng.normal(size=3) * 10.0 ** rng.integers(-3, 4)
                toks = [(fmts[(f + a + d) % 4] % v[d]) for d in range(3)]
                if (f + a) % 5 == 0:
                    toks = [t.replace("e", "D") for t in toks]
                sep = "\t" if a % 3 == 0 else "  "
                eol = "\r\n" if f % 4 == 0 else "\n"
                last = (f == nfr - 1 and a == nat - 1)
                fh.write(("%5d%sX%d%s" % (a + 1, sep, a, sep)) + sep.join(toks) + "  7 8 9" + ("" if last else eol))
    printed = np.array([2, 3, 7, 11, 23])
    sym_n, got = parseMD._read_blocks(str(path), nat, 2, (1, 2, 3, 4), printed, 0.529177)
    sym_p, ref = parseMD._read_blocks_py(str(path), nat, 2, (1, 2, 3, 4), printed)
    assert sym_n == sym_p
    np.testing.assert_array_equal(got.numpy(), ref / 0.529177)
    # errors are loud: file too short, malformed number
    with pytest.raises(_lib.DynpyError, match="fewer than"):
        parseMD._read_blocks(str(path), nat, 2, (1, 2, 3, 4), np.array([24]))
    bad = tmp_path / "bad.xyz"
    bad.write_text("1\nc\nH 0.0 abc 1.0\n")
    with pytest.raises(_lib.DynpyError, match="cannot parse 'abc'"):
        parseMD._read_blocks(str(bad), 1, 2, (0, 1, 2, 3), np.array([1]))


def test_parse_qe_pos_and_vel(tmp_path):
    """MD_format = 'QE' (parseMD.py:67-153): the directory is `traj` itself (not traj_dir + traj), frame ids are the
    step numbers of the kept .pos header lines, coordinates are not converted, symbols come from the parameter
    class; .vel values sit in columns 1-3.  The committed fixture the reference ran (sr_qe) is parsed too."""
    from dynpy_b200 import parseMD
    rng = np.random.default_rng(9)
    nat, nfr = 5, 12
    pos = rng.normal(size=(nfr, nat, 3)) * 7
    vel = rng.normal(size=(nfr, nat, 3))
    d = tmp_path / "run7"
    d.mkdir()
    with open(d / "h2o.pos", "w") as fp, open(d / "h2o.vel", "w") as fv:
        for f in range(nfr):
            fp.write("  %d   %.6f\n" % (50 * (f + 1), 0.0121 * (f + 1)))
            fv.write("  %d   %.6f\n" % (50 * (f + 1), 0.0121 * (f + 1)))
            for a in range(nat):
                fp.write("  %.17g %.17g %.17g\n" % tuple(pos[f, a]))
                fv.write("H  %.17g %.17g %.17g\n" % tuple(vel[f, a]))

    class PD:
        MD_format = "QE"; traj_dir = "/nonexistent/"; trajs = [str(d)]; nat = 5; start_prod = 3; end_prod = 11
        md_print_freq = 1000; sample_freq = 4; celldm = 20.0; timestep = 0.001; parse_vel = True
        symbols = ["O", "H", "H", "Na", "Cl"]
    tr = parseMD.PARSE_MD(PD, str(d), device=torch.device("cpu"))
    assert tr.symbols == PD.symbols
    np.testing.assert_array_equal(tr.frames, [150, 350, 550])          # header step ids, no md_print_freq
    np.testing.assert_array_equal(tr.pos.numpy(), pos[[2, 6, 10]].transpose(1, 2, 0))   # bit-exact, no unit change
    np.testing.assert_array_equal(tr.vel.numpy(), vel[[2, 6, 10]].transpose(1, 2, 0))
    # the fixture the unmodified reference ran: 6 waters, printed frames 2..26 step 2, header ids 10 x printed
    cwd = os.getcwd()
    os.chdir(os.path.join(GOLDEN, "sr_qe"))
    try:
        P = load_params("sr_qe").ParseDynamics
        tr = parseMD.PARSE_MD(P, "01", device=torch.device("cpu"))
    finally:
        os.chdir(cwd)
    np.testing.assert_array_equal(tr.frames, np.arange(2, 27, 2) * 10)
    assert tr.pos.shape == (18, 3, 13) and tr.symbols == P.symbols

    class Bad(PD):
        symbols = ["O", "H"]
    with pytest.raises(ValueError, match="symbols lists 2 atoms"):
        parseMD.PARSE_MD(Bad, str(d), device=torch.device("cpu"))


# ------------------------------------------------------------------ topology / pair indexing
def test_topology_numbering_and_pair_selection_match_oracle():
    from dynpy_b200 import DDparse, synth, topology
    box = synth.water_box(8, seed=2)
    pos = box.frames(0, 3).numpy().transpose(2, 0, 1)
    two, mol = orc.atom_two_table(pos, box.symbols, box.celldm, 15.0)
    f0 = two["f"] == 0
    b = two["bond"] & f0
    topo = topology.Topology(box.nat, two["i"][b], two["j"][b], box.symbols)
    np.testing.assert_array_equal(topo.mol_rank, mol[0])
    assert topo.nmol == 8 and topo.n_isolated == 0
    assert topo.molecule_id(2, 5, 3) == mol[2][topo.groups[5][0]]
    assert topo.classify_exact("H(2)O(1)").all()
    with pytest.raises(KeyError):
        topo.classify_exact("H(4)C(1)")
    # isolated atoms: numbered first, then overwritten by the connected_components walk (universe.py:426-437),
    # so all components are numbered in atom order starting at n_isolated (pinned by tests/golden/nn_water)
    t2 = topology.Topology(5, np.array([1]), np.array([3]), ["Na", "O", "Cl", "H", "K"])
    np.testing.assert_array_equal(t2.mol_rank, [0, 1, 2, 1, 3])
    assert t2.n_isolated == 3 and t2.nmol == 4
    np.testing.assert_array_equal(t2.frame_molecule_id(t2.mol_rank), [3, 4, 5, 4, 6])
    assert t2.molecule_id(1, 1, 10) == 10 * 3 + 1 * 4 + 1
    np.testing.assert_array_equal(orc._components(5, np.array([1]), np.array([3])), [3, 4, 5, 4, 6])
    for ptype, alabel in (("all", None), ("intra", None), ("inter", None), ("all", 1), ("all", 0)):
        pi, pj = DDparse.select_pairs(box.symbols, topo, 3, ptype, alabel)
        sym = np.array(box.symbols)
        i, j = two["i"][f0], two["j"][f0]
        if alabel:
            sel = ((i % 3) == alabel) & (sym[j] == "H")
        else:
            sel = (sym[i] == "H") & (sym[j] == "H")
        if ptype == "intra":
            sel &= mol[0][i] == mol[0][j]
        if ptype == "inter":
            sel &= mol[0][i] != mol[0][j]
        np.testing.assert_array_equal(pi, i[sel])
        np.testing.assert_array_equal(pj, j[sel])
    assert topology.atoms_per_formula("H(3)C(2)N(1)") == 6
    assert topology.parse_formula("H(2)O(1)") == {"H": 2, "O": 1}


def test_sr_frame_step_quirk():
    from dynpy_b200 import SRparse
    assert SRparse.frame_step(np.array([4, 6, 8, 10]), 5) == 2
    assert SRparse.frame_step(np.array([4, 6, 9]), 5) == 3          # last *new* distinct diff
    assert SRparse.frame_step(np.array([4, 6, 9, 11]), 5) == 3      # 2 was already seen
    assert SRparse.frame_step(np.array([4, 6]), 1) == 2


def test_simpson_mode_default_follows_scipy():
    from dynpy_b200 import config
    import scipy
    maj, mnr = (int(x) for x in scipy.__version__.split(".")[:2])
    assert config.default_simpson_mode() == ("cartwright" if (maj, mnr) >= (1, 11) else "avg")
    os.environ["DYNPY_SIMPSON"] = "avg"
    try:
        assert config.default_simpson_mode() == "avg"
    finally:
        del os.environ["DYNPY_SIMPSON"]


# ------------------------------------------------------------------ multi-rank plumbing
def test_shard_range_partitions_everything():
    from dynpy_b200 import dist
    for n in (0, 1, 2, 7, 130816, 523264):
        for world in (1, 2, 3, 8):
            cuts = [dist.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts[:-1], cuts[1:]):
                assert a[1] == b[0] and a[0] % 2 == 0
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 2


_WORKER = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, %(root)r)
from dynpy_b200 import dist
from oracle import oracle_np as orc
rank, world, local = dist.init_from_env("gloo")
rng = np.random.default_rng(0)
x = rng.normal(size=(11, 64))
lo, hi = dist.shard_range(11)
# each rank sums the ACFs of its shard (CPU stand-in for the per-rank spectrum), then ONE all-reduce
part = np.zeros(64)
for s in range(lo, hi):
    part += orc.wiener_khinchin(x[s])
t = torch.from_numpy(part)
dist.allreduce_sum_(t)
ref = sum(orc.wiener_khinchin(x[s]) for s in range(11))
assert np.allclose(t.numpy(), ref, rtol=0, atol=1e-12), "rank %%d mismatch" %% rank
dist.barrier()
print("rank", rank, "of", world, "shard", lo, hi, "ok")
"""


def test_two_rank_gloo_allreduce(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT})
    port = 29000 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2, r.stdout


def test_frame_topologies_match_per_frame_molecule_detection():
    """A trajectory whose bond list changes in one frame (two waters touch): FrameTopologies (static topology +
    the deviating frames) must give the molecule numbers of the reference's all-frames compute_molecule, and
    select_molecules must drop the molecules that are not intact in every frame (SRparse.py:171-176)."""
    from dynpy_b200 import SRparse, synth, topology
    box = synth.water_box(6, seed=21)
    pos = box.frames(0, 26).numpy().transpose(2, 0, 1)[1::2]
    F, A, _ = pos.shape
    two, mol = orc.atom_two_table(pos, box.symbols, box.celldm, 8.0)
    per_frame = []
    for f in range(F):
        b = two["bond"] & (two["f"] == f)
        per_frame.append(set(zip(two["i"][b].tolist(), two["j"][b].tolist())))
    changes = {f: sorted(per_frame[f] ^ per_frame[0]) for f in range(F) if per_frame[f] != per_frame[0]}
    assert changes, "the fixture is meant to have a frame with an extra contact"
    b0 = sorted(per_frame[0])
    base = topology.Topology(A, np.array([x[0] for x in b0]), np.array([x[1] for x in b0]), box.symbols)
    ft = topology.FrameTopologies(base, changes, F)
    np.testing.assert_array_equal(ft.molecule_ids(np.arange(A)), mol)

    class SR:
        mol_type = "water"; nmol = 6
    keep, nat = SRparse.select_molecules(ft, SR)
    touched = {a for pairs in changes.values() for ij in pairs for a in ij}
    want = [m for m in range(6) if not (touched & set(range(3 * m, 3 * m + 3)))]
    assert [k[0] for k in keep] == want and nat == 3
    assert [k[2] for k in keep] == [(3 * m, 3 * m + 1, 3 * m + 2) for m in want]


def test_qr_missing_values_follow_reference_complex_arithmetic():
    """qrax._fill_missing (host-side input cleaning of the QR driver): the filled Cartesian set must give the
    same five spherical series as interpolating Re/Im of the reference's R_2m — including `1j * NaN` voiding the
    real part (fixture qr_nan, whose ACFs the oracle reproduces from the reference's output)."""
    from dynpy_b200 import qrax
    rows = orc.read_efg_csv(os.path.join(GOLDEN, "qr_nan", "efg.csv"))
    names = ("Vxx", "Vyy", "Vzz", "Vxy", "Vxz", "Vyz")
    for lab in (0, 1):
        m = rows["label"] == lab
        v = np.stack([rows[k][m] for k in names])
        assert np.isnan(v).any()
        f = qrax._fill_missing(v)
        assert not np.isnan(f).any()
        want = orc.qr_spatial({k: v[i] for i, k in enumerate(names)})
        want = [orc.interpolate_nan(np.real(r)) + 1j * orc.interpolate_nan(np.imag(r)) for r in want]
        got = orc.qr_spatial({k: f[i] for i, k in enumerate(names)})
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-15)
    clean = np.arange(12.0).reshape(6, 2)
    assert qrax._fill_missing(clean) is clean


def test_derived_topology_equals_full_reindexing():
    """Topology.derive (only the molecules that hold an end of a toggled bond are re-indexed) against a full
    union-find of the toggled bond list: random graphs with isolated atoms, bonds added (merges) and removed
    (splits, new isolated atoms)."""
    from dynpy_b200 import topology
    rng = np.random.default_rng(5)
    for trial in range(60):
        A = int(rng.integers(8, 60))
        nb = int(rng.integers(0, A))
        iu, ju = np.triu_indices(A, k=1)
        pick = rng.choice(len(iu), size=min(nb, len(iu)), replace=False)
        bonds = sorted(zip(iu[pick].tolist(), ju[pick].tolist()))
        sym = ["H" if a % 3 else "O" for a in range(A)]
        base = topology.Topology(A, np.array([b[0] for b in bonds], dtype=np.int64),
                                 np.array([b[1] for b in bonds], dtype=np.int64), sym)
        ntog = int(rng.integers(1, 5))
        cand = [bonds[k] for k in rng.choice(len(bonds), size=min(ntog, len(bonds)), replace=False)] if bonds else []
        extra = rng.choice(len(iu), size=ntog, replace=False)
        toggled = sorted(set(cand[:ntog // 2 + 1]) | {(int(iu[k]), int(ju[k])) for k in extra[:ntog // 2 + 1]})
        new = sorted(set(bonds) ^ set(toggled))
        full = topology.Topology(A, np.array([b[0] for b in new], dtype=np.int64),
                                 np.array([b[1] for b in new], dtype=np.int64), sym)
        der = base.derive(toggled)
        np.testing.assert_array_equal(der.mol_rank, full.mol_rank)
        assert (der.nmol, der.n_isolated) == (full.nmol, full.n_isolated)
        assert der.groups == full.groups
        np.testing.assert_array_equal(der.classify_exact({"O": 1, "H": 2}) if any(
            full.formula(m) == {"O": 1, "H": 2} for m in range(full.nmol)) else [], full.classify_exact(
            {"O": 1, "H": 2}) if any(full.formula(m) == {"O": 1, "H": 2} for m in range(full.nmol)) else [])


def test_deviating_frame_lists_group_lazily():
    """{frame: [(i, j)]} view over the device scan's sorted event rows (topology._LazyPairs): sizes and truth value
    without building per-frame lists, look-ups and iteration in frame order, empty scan."""
    from dynpy_b200 import topology
    ev = np.array([[2, 1, 5], [2, 3, 4], [7, 0, 1], [9, 2, 3], [9, 2, 4], [9, 5, 6]], dtype=np.int64)
    lp = topology._group_events(ev)
    assert len(lp) == 3 and bool(lp) and 7 in lp and 3 not in lp
    assert lp[2] == [(1, 5), (3, 4)] and lp[9] == [(2, 3), (2, 4), (5, 6)]
    assert list(lp.keys()) == [2, 7, 9] and list(lp) == [2, 7, 9]
    assert dict(lp.items()) == {2: [(1, 5), (3, 4)], 7: [(0, 1)], 9: [(2, 3), (2, 4), (5, 6)]}
    # rows that arrive unsorted (merged per-rank lists) are sorted by (frame, i, j) first
    lp2 = topology._group_events(ev[[3, 0, 5, 2, 1, 4]])
    assert dict(lp2.items()) == dict(lp.items())
    empty = topology._group_events(np.empty((0, 3), np.int64))
    assert len(empty) == 0 and not empty and list(empty.items()) == []
