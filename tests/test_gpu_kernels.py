"""GPU parity tests: every C-ABI entry point against the CPU oracle (oracle/oracle_np.py)
and the committed reference vectors.  Run on the B200 box:  pytest -m gpu.

Tolerances: ACF columns normwise max|dG|/max|G| <= 1e-9 (we assert much
tighter where the arithmetic allows), pair lists bit-exact, rates <= 1e-8 relative.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import oracle_np as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dynpy_b200 import ops as o
    return o


def expected_fft_length(n):
    """dynpy_fft_length: smallest of 2^k and 4096 x {20, 24, 40, 50} that is >= 2n - 1 (and >= 32)."""
    m = 32
    while m < 2 * n:
        m <<= 1
    for m1 in (20, 24, 40, 50):
        if m1 * 4096 >= 2 * n - 1 and m1 * 4096 < m:
            m = m1 * 4096
    return m


def dev(x, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).cuda()


def nerr(got, ref):
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)


# ------------------------------------------------------------------ engine: all plans
@pytest.mark.parametrize("N", [1, 5, 16, 17, 33, 100, 316, 700, 1025, 2048,       # single pass M = 32 .. 4096
                               2049, 5000, 9000, 20000, 36000, 45000, 60000, 70000, 100000, 120000,  # M1 = 2 .. 64 (20, 24, 40, 50 mixed radix)
                               150000, 300000, 600000])                          # M1 = 128 .. 512
def test_wk_sum_complex_all_plans(ops, N):
    rng = np.random.default_rng(N)
    S = 3 if N > 20000 else 5
    z = rng.normal(size=(S, N)) + 1j * rng.normal(size=(S, N))
    z *= np.exp(-np.arange(N) / (0.3 * N + 1))[None]
    spec = ops.wk_acf_sum(dev(z.real), dev(z.imag))
    M = spec.numel()
    assert M == expected_fft_length(N)
    acf = ops.ifft_acf(spec, N, 1.0 / (M * N)).cpu().numpy()[0]
    ref = sum(orc.wiener_khinchin(z[s]) for s in range(S))
    assert nerr(acf, ref) < 1e-12, nerr(acf, ref)


@pytest.mark.parametrize("N", [82000, 90000, 102400])
def test_wk_mixed_radix_length_matches_power_of_two(ops, N):
    """M = 50 x 4096 (radix-5 column pass) and M = 2^18 give the same linear ACF (any M >= 2N-1 does)."""
    rng = np.random.default_rng(N)
    z = rng.normal(size=(4, N))
    out = {}
    for M in (204800, 1 << 18):
        spec = ops.wk_acf_sum(dev(z), pack_real_pairs=True, M=M)
        assert spec.numel() == M
        out[M] = ops.ifft_acf(spec, N, 1.0 / (M * N)).cpu().numpy()[0]
    ref = sum(orc.wiener_khinchin(z[s]) for s in range(4))
    assert nerr(out[204800], ref) < 1e-12
    assert nerr(out[204800], out[1 << 18]) < 1e-12


@pytest.mark.parametrize("N", [1 << 20, 3000000, 6000000])  # M1 = 512, 2048, 4096 (largest supported)
def test_wk_sum_long_series(ops, N):
    rng = np.random.default_rng(7)
    z = rng.normal(size=(2, N))
    spec = ops.wk_acf_sum(dev(z), pack_real_pairs=True)
    M = spec.numel()
    acf = ops.ifft_acf(spec, N, 1.0 / (M * N)).cpu().numpy()[0]
    ref = orc.wiener_khinchin(z[0]) + orc.wiener_khinchin(z[1])
    assert nerr(acf, ref) < 1e-12


@pytest.mark.parametrize("S,N", [(1, 50), (2, 50), (7, 333), (9, 5000), (3, 70000)])
def test_wk_real_pair_packing_and_scale(ops, S, N):
    rng = np.random.default_rng(S * 1000 + N)
    x = rng.normal(size=(S, N)).cumsum(axis=1) / np.sqrt(N)
    spec = ops.wk_acf_sum(dev(x), pack_real_pairs=True)
    M = spec.numel()
    acf = ops.ifft_acf(spec, N, 1.0 / (M * N)).cpu().numpy()[0]
    ref = sum(orc.wiener_khinchin(x[s]) for s in range(S))
    assert nerr(acf, ref) < 1e-12
    w = rng.uniform(0.5, 2.0, S)
    spec = ops.wk_acf_sum(dev(x), scale=dev(w))
    acf = ops.ifft_acf(spec, N, 1.0 / (M * N)).cpu().numpy()[0]
    ref = sum(w[s] ** 2 * orc.wiener_khinchin(x[s]) for s in range(S))
    assert nerr(acf, ref) < 1e-12


def test_wk_accumulates_and_matches_reference_vector(ops):
    d = np.load(os.path.join(GOLDEN, "leaf_vectors.npz"))
    z = d["wk_in"]
    spec = ops.wk_acf_sum(dev(z.real[None]), dev(z.imag[None]))
    spec = ops.wk_acf_sum(dev(z.real[None]), dev(z.imag[None]), spectrum=spec)  # second call accumulates
    M, N = spec.numel(), len(z)
    acf = ops.ifft_acf(spec, N, 0.5 / (M * N)).cpu().numpy()[0]
    assert nerr(acf, d["wk_out"]) < 1e-13


def test_argument_errors_are_loud(ops):
    from dynpy_b200._lib import DynpyError
    with pytest.raises(DynpyError):
        ops.wk_acf_sum(torch.zeros((2, 8), dtype=torch.float64))  # CPU tensor: no CPU path
    x = torch.zeros((2, 100), dtype=torch.float64, device="cuda")
    with pytest.raises(DynpyError):
        ops.wk_acf_sum(x, M=128)  # M < 2N-1
    with pytest.raises(DynpyError):
        ops.wk_acf_sum(x, M=384)  # not a power of two


# ------------------------------------------------------------------ QR
@pytest.mark.parametrize("S,N", [(1, 64), (2, 81), (5, 316), (4, 4096), (3, 70000)])
def test_qr_spectrum(ops, S, N):
    from dynpy_b200 import synth
    efg = synth.efg_series(S, N, seed=S + N)
    spec = ops.qr_efg_to_spectrum(efg.cuda())
    M = spec.shape[1]
    acf = ops.ifft_acf(spec, N, 1.0 / (M * N)).cpu().numpy()
    e = efg.numpy()
    ref = np.zeros((5, N))
    for s in range(S):
        V = dict(Vxx=e[s, 0], Vyy=e[s, 1], Vzz=e[s, 2], Vxy=e[s, 3], Vxz=e[s, 4], Vyz=e[s, 5])
        for m, r in enumerate(orc.qr_spatial(V)):
            ref[m] += orc.wiener_khinchin(r)
    for m, a in ((0, 2), (1, 1), (2, 0), (3, 1), (4, 2)):  # m = -2..2  <-  acc (R22, R21, R20)
        assert nerr(acf[a], ref[m]) < 1e-12


# ------------------------------------------------------------------ wrap + pdist (bit-exact)
def test_wrap_bit_exact(ops):
    rng = np.random.default_rng(3)
    L = 30.651558
    k = np.arange(-60, 61) * L  # near-integer quotients: the fast exact-fmod path must fix q by one
    x = np.concatenate([rng.uniform(-3 * L, 3 * L, 5000), [0.0, -0.0, L, -L, -1e-18, 1e-18, 2 * L, -1e-300, 5e-324],
                        k, np.nextafter(k, np.inf), np.nextafter(k, -np.inf), rng.uniform(-1e7 * L, 1e7 * L, 2000),
                        rng.uniform(-1e11 * L, 1e11 * L, 500), rng.uniform(-1e15 * L, 1e15 * L, 500), [1e300, -1e300]])
    got = ops.wrap(dev(x), L).cpu().numpy()
    np.testing.assert_array_equal(got, np.mod(x, L))
    assert got.max() <= L


def test_pdist_reference_vectors(ops):
    d = np.load(os.path.join(GOLDEN, "leaf_vectors.npz"))
    for k in range(4):
        x, y, z = d["pd%d_in" % k]
        L, dmax = d["pd%d_par" % k]
        got = ops.pdist_ortho(dev(x), dev(y), dev(z), L, L, L, dev(d["pd%d_idx" % k], torch.int64), dmax)
        for c, nm in enumerate(("dx", "dy", "dz", "dr", "a0", "a1", "prj")):
            np.testing.assert_array_equal(got[c].cpu().numpy(), d["pd%d_%s" % (k, nm)], err_msg="%d %s" % (k, nm))


@pytest.mark.parametrize("n,dmax", [(0, 3.0), (1, 3.0), (2, 3.0), (2, 0.01), (300, 4.0), (768, 15.0), (1000, 1e9)])
def test_pdist_vs_oracle(ops, n, dmax):
    rng = np.random.default_rng(n)
    a, b, c = 17.3, 19.1, 23.7
    x, y, z = rng.uniform(0, a, n), rng.uniform(0, b, n), rng.uniform(0, c, n)
    idx = np.arange(n, dtype=np.int64) * 3 + 7
    got = ops.pdist_ortho(dev(x), dev(y), dev(z), a, b, c, dev(idx, torch.int64), dmax)
    ref = orc.pdist_ortho(x, y, z, a, b, c, idx, dmax) if n >= 2 else [np.zeros(0)] * 7
    for k in range(7):
        np.testing.assert_array_equal(got[k].cpu().numpy(), ref[k])


def test_pdist_strided_frame_layout(ops):
    rng = np.random.default_rng(9)
    A, F, L = 40, 6, 11.0
    X = rng.uniform(0, L, (A, 3, F))
    Xd = dev(X)
    f = 4
    got = ops.pdist_ortho(Xd[:, 0, f], Xd[:, 1, f], Xd[:, 2, f], L, L, L, dev(np.arange(A), torch.int64), 5.0,
                          stride=3 * F)
    ref = orc.pdist_ortho(X[:, 0, f], X[:, 1, f], X[:, 2, f], L, L, L, np.arange(A), 5.0)
    for k in range(7):
        np.testing.assert_array_equal(got[k].cpu().numpy(), ref[k])


# ------------------------------------------------------------------ Simpson
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 10, 11, 316, 317, 100000, 100001])
@pytest.mark.parametrize("mode", ["cartwright", "avg"])
def test_simpson(ops, n, mode):
    rng = np.random.default_rng(n)
    x = np.cumsum(rng.uniform(0.5, 1.5, n)) if n < 1000 else np.arange(n) * 0.001
    y = rng.normal(size=(3, n)) + np.exp(-np.arange(n) / (0.2 * n + 1))
    got = ops.simpson(dev(y), dev(x), mode).cpu().numpy()
    ref = np.array([orc.simpson(y[c], x, mode) for c in range(3)])
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)


# ------------------------------------------------------------------ dipolar fast path (dd2): every column plan
@pytest.mark.parametrize("F,P,budget_mb", [
    (5000, 7, 64),        # M = 16384  (M1 = 4: generic engine, not the fast path)
    (20000, 9, 64),       # M = 65536  (M1 = 16), odd pair count
    (36000, 5, 4096),     # M = 81920  (M1 = 20, mixed radix 2 x 10)
    (45000, 4, 4096),     # M = 98304  (M1 = 24, mixed radix 2 x 12)
    (60000, 6, 4096),     # M = 131072 (M1 = 32)
    (70000, 5, 200),      # M = 163840 (M1 = 40, mixed radix 2 x 20), odd pair count, several batches
    (90000, 5, 200),      # M = 204800 (M1 = 50, mixed radix 2 x 25), odd pair count, several batches
    (102400, 4, 4096),    # M = 204800 at its largest N (all 25 x 4096 frames in use)
    (110000, 3, 4096),    # M = 262144 (M1 = 64)
])
def test_dd_fast_path_vs_oracle(ops, F, P, budget_mb):
    """coords -> accumulated spectra -> summed ACFs of the fused dipolar path against the numpy oracle
    (27-image minimum image, angle-form Y2m, one fft/ifft per series): normwise <= 1e-10 per column."""
    from dynpy_b200 import synth
    box = synth.water_box(4, seed=21, drift=0.01, trans_amp=0.2).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    raw = box.trajectory(F, chunk=4096, atom_mask=hidx)
    L = box.celldm
    w = ops.wrap(raw, L)
    i, j = np.triu_indices(8, k=1)
    i, j = i[:P], j[:P]
    spec = ops.dd_pairs_to_spectrum(w, dev(i, torch.int32), dev(j, torch.int32), (L, L, L), ws_budget=budget_mb << 20)
    M = spec.shape[1]
    assert M == ops.fft_length(F)
    G = ops.ifft_acf(spec, F, 1.0 / (float(M) * F)).cpu().numpy()
    ref = orc.dd_pairs_gapless(w.cpu().numpy(), list(zip(i.tolist(), j.tolist())), L)
    scale = max(np.max(np.abs(ref[c])) for c in (4, 5, 6))
    for c in range(7):
        denom = np.max(np.abs(ref[c])) if c != 3 else max(np.max(np.abs(ref[3])), 1e-6 * scale)
        assert np.max(np.abs(G[c] - ref[c])) / denom < 1e-10, (c, np.max(np.abs(G[c] - ref[c])) / denom)


@pytest.mark.parametrize("F,P", [(100000, 301), (600000, 23)])
def test_dd_fast_path_matches_generic_engine(ops, monkeypatch, F, P):
    """DYNPY_DD_PATH=v1 routes the same call through the generic word generator + engine: both paths
    must agree to rounding on a workload large enough for many batches and work ranges (1e5 frames: the
    thread-per-column kernel, M1 = 50; 6e5 frames: the two-level kernel for 512-point columns, M = 2^21)."""
    from dynpy_b200 import synth
    box = synth.water_box(16, seed=3).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    w = ops.wrap(box.trajectory(F, chunk=1 << 15, atom_mask=hidx), box.celldm)
    i, j = np.triu_indices(32, k=1)
    pi, pj = dev(i[:P], torch.int32), dev(j[:P], torch.int32)
    L = box.celldm
    out = {}
    for path in ("v1", "v2"):
        monkeypatch.setenv("DYNPY_DD_PATH", path)
        spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), ws_budget=1 << 30)
        out[path] = ops.ifft_acf(spec, F, 1.0 / (float(spec.shape[1]) * F)).cpu().numpy()
    for c in range(7):
        assert nerr(out["v2"][c], out["v1"][c]) < 1e-12


@pytest.mark.parametrize("F", [100000, 99998, 1000000])
def test_dd_geometry_kernel_variants_agree(ops, monkeypatch, F):
    """The three forms of the pair-geometry kernel (scalar loads, 32-byte loads, TMA bulk copies through an
    mbarrier ring) feed the same arithmetic: identical spectra up to the summation order of the r^-3 mean.
    99998 frames: N % 4 != 0 (the 32-byte form falls back to scalar loads, the TMA form still runs)."""
    from dynpy_b200 import synth
    box = synth.water_box(8, seed=3).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=1 << 15, atom_mask=hidx), L)
    i, j = np.triu_indices(16, k=1)
    P = 37 if F <= 100000 else 5
    pi, pj = dev(i[:P], torch.int32), dev(j[:P], torch.int32)
    out = {}
    for mode in ("scalar", "vec", "tma"):
        monkeypatch.setenv("DYNPY_DD_GEOM", mode)
        out[mode] = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), ws_budget=1 << 30)
    scale = out["scalar"].abs().amax(dim=1, keepdim=True)
    for mode in ("vec", "tma"):
        d = ((out[mode] - out["scalar"]).abs() / scale).amax(dim=1).cpu().numpy()
        assert d.max() < 1e-12, (mode, d)   # same per-pair-frame arithmetic; the accumulation order of the spectra
                                            # (atomic adds) and of the r^-3 mean differs from run to run


def test_dd_full_size_properties(ops):
    """cfg2 frame count (1e5 frames, M = 204800), thousands of pairs: size-independent properties of the fused
    dipolar path — (1) the accumulated spectrum is additive over disjoint pair sets whatever the batching,
    (2) lag 0 of every summed ACF equals the directly computed mean square of its series (Parseval), with the
    series formed by independent torch code (27-image-free nearest image via rounding, angle-free Y2m)."""
    from dynpy_b200 import synth
    F = 100000
    box = synth.water_box(32, seed=2).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=4096, atom_mask=hidx), L)
    i, j = np.triu_indices(64, k=1)
    i, j = i[:1501], j[:1501]           # odd count: the last pair-pair has one member
    pi, pj = dev(i, torch.int32), dev(j, torch.int32)
    cell = (L, L, L)
    whole = ops.dd_pairs_to_spectrum(w, pi, pj, cell, ws_budget=4 << 30)
    M = whole.shape[1]
    assert M == 204800
    part = ops.dd_pairs_to_spectrum(w, pi[:700], pj[:700], cell, ws_budget=300 << 20)
    part = ops.dd_pairs_to_spectrum(w, pi[700:], pj[700:], cell, spectrum=part, ws_budget=1 << 30)
    scale = whole.abs().amax(dim=1, keepdim=True)
    assert float(((whole - part).abs() / scale).max()) < 1e-12
    G0 = ops.ifft_acf(whole, F, 1.0 / (float(M) * F))[:, 0].cpu().numpy()
    # independent evaluation of sum_pairs mean_t |series|^2
    ref = np.zeros(7)
    for s0 in range(0, len(i), 256):
        a = w[torch.as_tensor(i[s0:s0 + 256], device="cuda")]
        b = w[torch.as_tensor(j[s0:s0 + 256], device="cuda")]
        d = a - b
        d = d - L * torch.round(d / L)
        r2 = (d * d).sum(dim=1)
        r3 = r2 ** -1.5
        x, y, z = d[:, 0], d[:, 1], d[:, 2]
        y20 = 0.25 * np.sqrt(5 / np.pi) * (3 * z * z / r2 - 1)
        y21sq = (0.5 * np.sqrt(15 / (2 * np.pi))) ** 2 * z * z * (x * x + y * y) / r2 ** 2
        y22sq = (0.25 * np.sqrt(15 / (2 * np.pi))) ** 2 * (x * x + y * y) ** 2 / r2 ** 2
        kf = 4 * np.pi / 5
        dr3 = r3 - r3.mean(dim=1, keepdim=True)
        for c, v in enumerate((y20 ** 2, y21sq, y22sq, dr3 ** 2, kf * r3 ** 2 * y20 ** 2, kf * r3 ** 2 * y21sq,
                               kf * r3 ** 2 * y22sq)):
            ref[c] += float(v.mean(dim=1).sum())
    np.testing.assert_allclose(G0, ref, rtol=1e-10)


@pytest.mark.parametrize("F,P,m1", [
    (1_000_000, 5, 512),   # cfg5 frame count: M = 2^21, the long-column kernel, odd pair count
    (600_000, 3, 512),     # M = 2^21 with the second half of the 512-point columns partly empty
    (140_000, 3, 128),     # M = 2^19
    (270_000, 3, 256),     # M = 2^20
])
def test_dd_long_series_full_acf_vs_oracle(ops, F, P, m1):
    """FULL summed ACFs (every lag, all 7 columns) of the fused dipolar call at the long transform lengths
    (M1 = 128, 256, 512; cfg5 is 1e6 frames = M 2^21) against orc.dd_pairs_gapless: normwise <= 1e-10."""
    from dynpy_b200 import synth
    box = synth.water_box(4, seed=5, drift=0.01, trans_amp=0.2).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=1 << 16, atom_mask=hidx), L)
    i, j = np.triu_indices(8, k=1)
    sel = [0, 9, 13, 20, 27][:P]          # an intramolecular pair and intermolecular ones
    i, j = i[sel], j[sel]
    M = ops.fft_length(F)
    assert M == m1 * 4096
    spec = ops.dd_pairs_to_spectrum(w, dev(i, torch.int32), dev(j, torch.int32), (L, L, L), M=M)
    G = ops.ifft_acf(spec, F, 1.0 / (float(M) * F)).cpu().numpy()
    ref = orc.dd_pairs_gapless(w.cpu().numpy(), list(zip(i.tolist(), j.tolist())), L)
    scale = max(np.max(np.abs(ref[c])) for c in (4, 5, 6))
    for c in range(7):
        denom = np.max(np.abs(ref[c])) if c != 3 else max(np.max(np.abs(ref[3])), 1e-6 * scale)
        err = np.max(np.abs(G[c] - ref[c])) / denom
        assert err < 1e-10, (c, err)


def test_dd_ragged_more_than_9362_pairs_per_batch(ops):
    """Short trajectory, normal-size box, rmax = 15 bohr: one ragged batch holds > 9362 partial pairs
    (65535 / 7 series: the old gridDim.y limit of the scatter launch).  Checked against the same call cut
    into small batches, and against the oracle on a subset of the pairs."""
    from dynpy_b200 import synth
    F = 48
    box = synth.water_box(96, seed=7, drift=0.02).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=64, atom_mask=hidx), L)
    i, j = np.triu_indices(192, k=1)
    keep = np.arange(len(i)) % 7 != 3
    i, j = i[keep][:11000], j[keep][:11000]
    assert len(i) > 9362
    pi, pj = dev(i, torch.int32), dev(j, torch.int32)
    cell = (L, L, L)
    big = ops.dd_pairs_ragged_to_acf(w, pi, pj, cell, 15.0, ws_budget=8 << 30)
    small = torch.zeros_like(big)
    for s0 in range(0, len(i), 3000):
        ops.dd_pairs_ragged_to_acf(w, pi[s0:s0 + 3000].contiguous(), pj[s0:s0 + 3000].contiguous(), cell, 15.0, G=small,
                                   ws_budget=64 << 20)
    scale = big.abs().amax(dim=1, keepdim=True)
    assert float(((big - small).abs() / scale).max()) < 1e-12
    sub = slice(5000, 5040)
    got = ops.dd_pairs_ragged_to_acf(w, pi[sub].contiguous(), pj[sub].contiguous(), cell, 15.0).cpu().numpy()
    ref = orc.dd_pairs_ragged(w.cpu().numpy(), list(zip(i[sub].tolist(), j[sub].tolist())), L, 15.0)
    scale = max(np.max(np.abs(ref[c])) for c in (4, 5, 6))
    for c in range(7):
        denom = np.max(np.abs(ref[c])) if c != 3 else max(np.max(np.abs(ref[3])), 1e-6 * scale)
        assert np.max(np.abs(got[c] - ref[c])) / denom < 1e-10, c


@pytest.mark.parametrize("npairs", [3, 4])
def test_dd_ragged_distant_pairs_keep_their_precision(ops, npairs):
    """256-water box, rmax = 15 bohr (the reference's notebooks/dynpy_params_water_dipolar.py), 30 000 frames: pairs
    that sit near rmax have r^-3-weighted ACFs ~1e-7 of their plain-harmonic ACFs.  The inverse transforms of the
    ragged path carry two even spectra each; packed without regard to magnitude the weak columns lost 7 digits
    (3e-8, caught by bench.py's gate).  Every column of a few such pairs against the oracle, odd and even counts."""
    from dynpy_b200 import synth
    F = 30000
    box = synth.water_box(256, seed=2).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=4096, atom_mask=hidx), L)
    i, j = np.triu_indices(512, k=1)
    i, j = i[::41][:3000], j[::41][:3000]
    pi, pj = dev(i, torch.int32), dev(j, torch.int32)
    cell = (L, L, L)
    counts = ops.dd_pair_count(w, pi, pj, cell, 15.0).cpu().numpy()
    part = np.nonzero((counts > 0) & (counts < F))[0]
    assert len(part) >= npairs
    sub = part[np.linspace(0, len(part) - 1, npairs).astype(int)]
    got = ops.dd_pairs_ragged_to_acf(w, pi[sub].contiguous(), pj[sub].contiguous(), cell, 15.0).cpu().numpy()
    atoms = np.unique(np.concatenate([i[sub], j[sub]]))
    remap = {int(a): k for k, a in enumerate(atoms)}
    ref = orc.dd_pairs_ragged(w[torch.from_numpy(atoms).cuda()].cpu().numpy(),
                              [(remap[int(a)], remap[int(b)]) for a, b in zip(i[sub], j[sub])], L, 15.0)
    scale = max(np.max(np.abs(ref[c])) for c in (4, 5, 6))
    for c in range(7):
        denom = np.max(np.abs(ref[c])) if c != 3 else max(np.max(np.abs(ref[3])), 1e-6 * scale)
        assert np.max(np.abs(got[c] - ref[c])) / denom < 1e-10, (c, np.max(np.abs(got[c] - ref[c])) / denom)


def test_dd_cfg5_frame_count_properties(ops):
    """cfg5 frame count (1e6 frames, M = 2^21, 512-point column transforms): lag 0 and a few scattered lags of
    the summed ACFs against direct O(N) sums of independently formed series, and additivity over pair sets."""
    from dynpy_b200 import synth
    F = 1_000_000
    box = synth.water_box(8, seed=5).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=1 << 16, atom_mask=hidx), L)
    i, j = np.triu_indices(16, k=1)
    i, j = i[:9], j[:9]
    pi, pj = dev(i, torch.int32), dev(j, torch.int32)
    cell = (L, L, L)
    M = ops.fft_length(F)
    assert M == 1 << 21
    whole = ops.dd_pairs_to_spectrum(w, pi, pj, cell, M=M)
    part = ops.dd_pairs_to_spectrum(w, pi[:4], pj[:4], cell, M=M)
    part = ops.dd_pairs_to_spectrum(w, pi[4:], pj[4:], cell, M=M, spectrum=part)
    scale = whole.abs().amax(dim=1, keepdim=True)
    assert float(((whole - part).abs() / scale).max()) < 1e-12
    G = ops.ifft_acf(whole, F, 1.0 / (float(M) * F))
    a = w[torch.as_tensor(i, device="cuda")]
    b = w[torch.as_tensor(j, device="cuda")]
    d = a - b
    d = d - L * torch.round(d / L)
    r2 = (d * d).sum(dim=1)
    r3 = r2 ** -1.5
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    kf = np.sqrt(4 * np.pi / 5)
    y20 = (0.25 * np.sqrt(5 / np.pi) * (3 * z * z / r2 - 1)).to(torch.complex128)
    y21 = -0.5 * np.sqrt(15 / (2 * np.pi)) * z * torch.complex(x, y) / r2
    y22 = 0.25 * np.sqrt(15 / (2 * np.pi)) * torch.complex(x, y) ** 2 / r2
    dr3 = (r3 - r3.mean(dim=1, keepdim=True)).to(torch.complex128)
    series = [y20, y21, y22, dr3, kf * r3 * y20, kf * r3 * y21, kf * r3 * y22]
    for lag in (0, 1, 777, 500_000, 999_999):
        for c, sr in enumerate(series):
            ref = float((sr[:, :F - lag].conj() * sr[:, lag:]).real.sum()) / F
            got = float(G[c, lag])
            assert abs(got - ref) <= 1e-10 * float(G[c].abs().max()), (lag, c, got, ref)


# ---------------------------------------------------------------- J(omega) / cutoff
@pytest.mark.parametrize("n", [1, 2, 3, 5, 60, 316, 777, 1000, 1009, 4096, 65536, 99991, 100000])
def test_dft_any_length_vs_numpy(ops, n):
    rng = np.random.default_rng(n)
    re = rng.normal(size=(2, n))
    im = rng.normal(size=(2, n))
    got = ops.dft(dev(re), dev(im)).cpu().numpy()
    ref = np.fft.fft(re + 1j * im, axis=1)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.0) * max(1.0, np.log2(n + 1))
    # real input, first half only (the shape ddrax.fourier_T asks for)
    got = ops.dft(dev(re[0]), n_out=n // 2).cpu().numpy()
    ref = np.fft.fft(re[0])[:n // 2]
    assert got.shape == (1, n // 2)
    if n >= 2:
        assert np.abs(got[0] - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.0) * max(1.0, np.log2(n + 1))


def test_dft_known_answers(ops):
    n = 1000
    j = np.arange(n)
    x = np.cos(2 * np.pi * 37 * j / n)
    X = ops.dft(dev(x)).cpu().numpy()[0]
    want = np.zeros(n, dtype=complex); want[37] = want[n - 37] = n / 2
    assert np.abs(X - want).max() < 1e-10
    X = ops.dft(dev(np.ones(997))).cpu().numpy()[0]      # prime length, constant series
    assert abs(X[0] - 997) < 1e-10 and np.abs(X[1:]).max() < 1e-10


def test_cutoff_bounds_vs_oracle(ops):
    g = np.load(os.path.join(GOLDEN, "specdens_vectors.npz"))
    for k in range(4):
        Y = g["Y%d" % k]
        for tol in (1e-3, 2e-2, 0.3):
            got = ops.cutoff_bounds(dev(Y), 100, tol).cpu().numpy()
            ref = np.array([orc.cutoff_bound(y, 100, tol) for y in Y])
            np.testing.assert_array_equal(got, ref)
    y = np.ones(300); y[150:] = 0.0
    assert ops.cutoff_bounds(dev(y)).tolist() == [249]
    assert ops.cutoff_bounds(dev(y), 100, 0.5).tolist() == [199]
    assert ops.cutoff_bounds(dev(y), 7, 0.0).tolist() == [156]
    rng = np.random.default_rng(3)
    Y = rng.normal(size=(6, 20000)) * np.exp(-np.arange(20000) / 900.0) + 1e-4
    Y[:, 0] = 1.0
    got = ops.cutoff_bounds(dev(Y), 100, 1e-3).cpu().numpy()
    np.testing.assert_array_equal(got, [orc.cutoff_bound(y) for y in Y])


def test_spectral_density_mirrors_vs_reference_vectors(ops):
    """ddrax.spec_dens / spec_dens_extr_narr(dt|cutoff) / dipolar(extr_narr=False) and
    qrax.spectral_dens(dt|cutoff) of the product against the reference's outputs (specdens_vectors.npz)."""
    from dynpy_b200 import ddrax, qrax
    g = np.load(os.path.join(GOLDEN, "specdens_vectors.npz"))
    for k in range(4):
        t, Y = g["t%d" % k], g["Y%d" % k]
        G = torch.zeros((7, len(t)), dtype=torch.float64, device="cuda")
        G[4:7] = dev(Y[:3])
        J = ddrax.spec_dens(dev(t), G)
        ref = g["J%d" % k]
        assert list(J.columns) == ddrax.J_COLUMNS
        np.testing.assert_allclose(J[ddrax.J_COLUMNS[0]].to_numpy(), ref[0].real, rtol=1e-15)
        np.testing.assert_allclose(J[ddrax.J_COLUMNS[2]].to_numpy(), ref[2].real, rtol=1e-15)
        np.testing.assert_allclose(J[ddrax.J_COLUMNS[1]].to_numpy()[1:], ref[1].real[1:], rtol=1e-15)
        for c in range(3):
            got = J[ddrax.J_COLUMNS[3 + c]].to_numpy()
            assert np.abs(got - ref[3 + c]).max() <= 1e-12 * np.abs(ref[3 + c]).max()
        for tag, kw in (("plain", {}), ("dt", dict(dt=0.002)), ("cut", dict(cutoff=True)),
                        ("cut2", dict(cutoff=True, cutoff_tol=2e-2))):
            want = g["dd_%s%d" % (tag, k)]
            e = ddrax.spec_dens_extr_narr(dev(t), G, "cartwright", **kw)
            got = [e['$j_{2,0}$'], e['$j_{2,1}$'], e['$j_{2,2}$'], e[r'$\tau_{c}$']]
            np.testing.assert_allclose(got, want, rtol=1e-10, err_msg="dd %s %d" % (tag, k))
            wq = g["qr_%s%d" % (tag, k)]
            q, qiso = qrax.spectral_dens(dev(t), dev(Y), "cartwright", **kw)
            np.testing.assert_allclose(q + [qiso], wq, rtol=1e-10, err_msg="qr %s %d" % (tag, k))
        if len(t) >= 100:
            Jd = dict(sqrt_omega=J[ddrax.J_COLUMNS[2]].to_numpy(), J0=J[ddrax.J_COLUMNS[3]].to_numpy(),
                      J1=J[ddrax.J_COLUMNS[4]].to_numpy(), J2=J[ddrax.J_COLUMNS[5]].to_numpy())
            for s2 in ("1H", "13C"):
                want = orc.dd_finite_freq(Jd, "1H", s2, larmor_freq=400)
                R, tc, f = ddrax.dipolar(None, "1H", s2, extr_narr=False, larmor_freq=400, J=J)
                assert tc is None and f is None
                np.testing.assert_allclose(R, want["R"], rtol=1e-12)
    with pytest.raises(NameError):
        ddrax.dipolar(None, "1H", "1H", extr_narr=False)
    one = ddrax.dipolar(2.5e-3, "1H", "1H", single_J=True)
    assert list(one.index) == ["1H"] and one.shape == (1, 1)


def test_wiener_khinchin_mirror_vs_reference_vectors(ops):
    """qrax.wiener_khinchin / ddrax.wiener_khinchin (the reference's call-site name) against outputs of the
    reference's own function: complex, complex with NaNs (interpolated per part), real; numpy and tensor input."""
    from dynpy_b200 import ddrax, qrax
    g = np.load(os.path.join(GOLDEN, "leaf_vectors.npz"))
    for tag, fn in (("wk", qrax.wiener_khinchin), ("wk_nan", qrax.wiener_khinchin), ("wk_real", ddrax.wiener_khinchin)):
        x, want = g[tag + "_in"], g[tag + "_out"]
        got = fn(x)
        assert isinstance(got, np.ndarray) and got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    x = g["wk_in"]
    t = torch.as_tensor(np.stack([x, 2 * x])).cuda()
    out = qrax.wiener_khinchin(t)
    assert out.is_cuda and out.shape == (2, len(x))
    assert np.abs(out[1].cpu().numpy() - 4 * g["wk_out"]).max() <= 1e-12 * 4 * np.abs(g["wk_out"]).max()


def test_cart_to_dipolar_from_df_vs_reference_vectors(ops):
    """ddrax.cart_to_dipolar_from_df (call-site signature) against the reference's numba ufuncs."""
    import pandas as pd
    from dynpy_b200 import ddrax
    g = np.load(os.path.join(GOLDEN, "leaf_vectors.npz"))
    d = g["dd_in"]
    df = pd.DataFrame({"dx": d[0], "dy": d[1], "dz": d[2], "dr": d[3], "time": np.arange(d.shape[1], dtype=float),
                       "label0": 0, "label1": 1})
    F = ddrax.cart_to_dipolar_from_df(df)
    assert list(F.columns) == ["time", "label0", "label1"] + ddrax.DIPOLAR_COLUMNS
    for col, key in zip(ddrax.DIPOLAR_COLUMNS, ("dd_r-3", "dd_Y20", "dd_Y21", "dd_Y22", "dd_F20", "dd_F21", "dd_F22")):
        want = g[key]
        got = F[col].to_numpy()
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), col


def test_wk_acf_sum_strided_rows(ops):
    """Series given as a strided view (component c of J[mol][3][frame]) must give exactly what a contiguous copy
    gives — the spin-rotation driver passes such views to avoid three copies of J."""
    torch.manual_seed(3)
    J = torch.randn((7, 3, 1500), dtype=torch.float64, device="cuda")
    for c in range(3):
        for pack in (True, False):
            a = ops.wk_acf_sum(J[:, c, :], pack_real_pairs=pack)
            b = ops.wk_acf_sum(J[:, c, :].contiguous(), pack_real_pairs=pack)
            assert torch.allclose(a, b, rtol=1e-12, atol=0.0)   # accumulation order differs between launches
    with pytest.raises(Exception):
        ops.wk_acf_sum(J[:, :, 5])      # series not contiguous
