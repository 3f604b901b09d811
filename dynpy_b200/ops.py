"""torch-tensor wrappers over the C-ABI (device pointers + the current CUDA stream).

PyTorch is plumbing here: it owns device memory and streams; every computation happens in
the hand-written kernels behind ``include/dynpy_b200.h``.  All tensors must be CUDA,
contiguous, float64 / int64 / int32 as documented; there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import os

import torch

from . import _lib

_WS = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.DynpyError("%s must be a CUDA tensor (dynpy_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise _lib.DynpyError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.DynpyError("%s must be contiguous" % name)
    return t


def _chk_rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D float64 CUDA tensor whose rows are contiguous (any row stride: a [:, c, :] view of [S, 3, N] works)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.DynpyError("%s must be a CUDA tensor (dynpy_b200 has no CPU path)" % name)
    if t.dtype != torch.float64 or t.dim() != 2:
        raise _lib.DynpyError("%s must be a 2-D float64 tensor" % name)
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise _lib.DynpyError("%s: the series (last axis) must be contiguous" % name)
    return t


def workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer (uint8), one per (device, current stream): kernels queued on a stream only ever
    share scratch with later kernels of the SAME stream, and a buffer that is replaced by a larger one is
    recorded on its stream first, so the caching allocator cannot hand its block to another stream while
    queued kernels still use it."""
    dev = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev)
    key = (dev, stream.cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        old = _WS.pop(key, None)
        if old is not None:
            old.record_stream(stream)
        del old, buf
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda:%d" % dev)
        _WS[key] = buf
    return buf


def release_workspace():
    _WS.clear()


def sm_count() -> int:
    n = _lib.load().dynpy_device_sm_count()
    if n < 0:
        _lib.check(n, "dynpy_device_sm_count")
    return n


def fft_length(n: int) -> int:
    return int(_lib.load().dynpy_fft_length(int(n)))


def _fft_ws(M: int, n_fft: int, device, budget: int = 2 << 30):
    lib = _lib.load()
    if os.environ.get("DYNPY_FFT_WS_MB"):      # A/B knob: batch size of the generic engine (T resident in L2 or not)
        budget = int(float(os.environ["DYNPY_FFT_WS_MB"]) * (1 << 20))
    one = lib.dynpy_fft_workspace_bytes(M, 1)
    if one == 0:
        return None, 0
    n = max(1, min(n_fft, budget // one))
    ws = workspace(device, n * one)
    return ws, ws.numel()


# ------------------------------------------------------------------------------ WK
def wk_acf_sum(re, im=None, pack_real_pairs=False, scale=None, M=None, spectrum=None):
    """spectrum[M] += sum_s |FFT_M(series_s)|^2 for series ``re[s] (+ i im[s])`` of shape [S, N]."""
    lib = _lib.load()
    re = _chk_rows(re, "re")
    S, N = re.shape
    stride = re.stride(0) if S > 1 else N
    if im is not None:
        im = _chk_rows(im, "im")
        assert im.shape == re.shape and (S == 1 or im.stride(0) == stride)
    if scale is not None:
        scale = _chk(scale, torch.float64, "scale")
    M = fft_length(N) if M is None else int(M)
    if spectrum is None:
        spectrum = torch.zeros(M, dtype=torch.float64, device=re.device)
    _chk(spectrum, torch.float64, "spectrum")
    with torch.cuda.device(re.device):
        n_fft = (S + 1) // 2 if pack_real_pairs else S
        ws, wsb = _fft_ws(M, n_fft, re.device)
        rc = lib.dynpy_wk_acf_sum_f64(re.data_ptr(), im.data_ptr() if im is not None else None, S, N, stride,
                                      1 if pack_real_pairs else 0, scale.data_ptr() if scale is not None else None,
                                      spectrum.data_ptr(), M, ws.data_ptr() if ws is not None else None, wsb, _stream())
    _lib.check(rc, "dynpy_wk_acf_sum_f64")
    return spectrum


def ifft_acf(spectrum, n_out: int, scale: float):
    """acf[a, k] = scale * Re(FFT_M(spectrum[a]))[k], k < n_out."""
    lib = _lib.load()
    spectrum = _chk(spectrum, torch.float64, "spectrum")
    if spectrum.dim() == 1:
        spectrum = spectrum[None]
    n_acc, M = spectrum.shape
    acf = torch.empty((n_acc, n_out), dtype=torch.float64, device=spectrum.device)
    with torch.cuda.device(spectrum.device):
        ws, wsb = _fft_ws(M, n_acc, spectrum.device)
        rc = lib.dynpy_ifft_acf_f64(spectrum.data_ptr(), n_acc, M, acf.data_ptr(), int(n_out), float(scale),
                                    ws.data_ptr() if ws is not None else None, wsb, _stream())
    _lib.check(rc, "dynpy_ifft_acf_f64")
    return acf


def wk_acf_each(re, im=None, M=None, ws_budget: int = 2 << 30):
    """Per-series wiener_khinchin: re (+ i im) [S, N] -> acf [S, N]."""
    lib = _lib.load()
    re = _chk(re, torch.float64, "re")
    S, N = re.shape
    if im is not None:
        im = _chk(im, torch.float64, "im")
    M = fft_length(N) if M is None else int(M)
    acf = torch.empty((S, N), dtype=torch.float64, device=re.device)
    one = int(lib.dynpy_wk_each_workspace_bytes(M, 1))
    n = max(1, min(S, ws_budget // one))
    with torch.cuda.device(re.device):
        wsb = int(lib.dynpy_wk_each_workspace_bytes(M, n))
        ws = workspace(re.device, wsb)
        rc = lib.dynpy_wk_acf_each_f64(re.data_ptr(), im.data_ptr() if im is not None else None, S, N, N,
                                       acf.data_ptr(), M, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "dynpy_wk_acf_each_f64")
    return acf


# ------------------------------------------------------------------------------ QR
def qr_efg_to_spectrum(efg, M=None, spectrum=None):
    """efg [S, 6, N] (Vxx,Vyy,Vzz,Vxy,Vxz,Vyz) -> spectrum [3, M] (R20, R2+-1, R2+-2), accumulated."""
    lib = _lib.load()
    efg = _chk(efg, torch.float64, "efg")
    S, six, N = efg.shape
    assert six == 6
    M = fft_length(N) if M is None else int(M)
    if spectrum is None:
        spectrum = torch.zeros((3, M), dtype=torch.float64, device=efg.device)
    _chk(spectrum, torch.float64, "spectrum")
    with torch.cuda.device(efg.device):
        ws, wsb = _fft_ws(M, 5 * ((S + 1) // 2), efg.device)
        rc = lib.dynpy_qr_efg_to_spectrum_f64(efg.data_ptr(), S, N, spectrum.data_ptr(), M,
                                              ws.data_ptr() if ws is not None else None, wsb, _stream())
    _lib.check(rc, "dynpy_qr_efg_to_spectrum_f64")
    return spectrum


# ------------------------------------------------------------------------------ DD
def wrap(x, L: float, out=None):
    lib = _lib.load()
    x = _chk(x, torch.float64, "x")
    out = torch.empty_like(x) if out is None else _chk(out, torch.float64, "out")
    with torch.cuda.device(x.device):
        rc = lib.dynpy_wrap_f64(x.data_ptr(), out.data_ptr(), x.numel(), float(L), _stream())
    _lib.check(rc, "dynpy_wrap_f64")
    return out


def pdist_ortho(ux, uy, uz, a, b, c, index, dmax=8.0, stride=1):
    """Bit-exact distances.pdist_ortho for one frame -> (dx, dy, dz, dr, atom0, atom1, projection).
    Large frames are counted first so that the outputs are sized by the rows found, not by n(n-1)/2."""
    lib = _lib.load()
    for t, nm in ((ux, "ux"), (uy, "uy"), (uz, "uz")):
        if not t.is_cuda or t.dtype != torch.float64:
            raise _lib.DynpyError("%s must be a CUDA float64 tensor" % nm)
    index = _chk(index, torch.int64, "index")
    n = index.numel()
    nn = n * (n - 1) // 2
    dev = index.device
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        wsb = int(lib.dynpy_pdist_workspace_bytes(n))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        args = (ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), int(stride), n, float(a), float(b), float(c),
                index.data_ptr(), float(dmax))
        cap = nn
        if nn > (1 << 20):
            rc = lib.dynpy_pdist_ortho_f64(*args, None, None, None, None, None, None, None, cnt.data_ptr(), ws.data_ptr(),
                                           wsb, _stream())
            _lib.check(rc, "dynpy_pdist_ortho_f64 (count)")
            cap = int(cnt.item())
        f = [torch.empty(max(cap, 1), dtype=torch.float64, device=dev) for _ in range(4)]
        g = [torch.empty(max(cap, 1), dtype=torch.int64, device=dev) for _ in range(3)]
        rc = lib.dynpy_pdist_ortho_f64(*args, f[0].data_ptr(), f[1].data_ptr(), f[2].data_ptr(), f[3].data_ptr(),
                                       g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                                       wsb, _stream())
    _lib.check(rc, "dynpy_pdist_ortho_f64")
    k = int(cnt.item())
    return tuple(t[:k] for t in f) + tuple(t[:k] for t in g)


def cart_to_dipolar(dx, dy, dz, dr):
    """Pair vectors [n] -> [11, n]: r^-3, Y20, Re/Im Y21, Re/Im Y22, F20, Re/Im F21, Re/Im F22."""
    lib = _lib.load()
    dx, dy, dz, dr = (_chk(t, torch.float64, nm) for t, nm in ((dx, "dx"), (dy, "dy"), (dz, "dz"), (dr, "dr")))
    n = dx.numel()
    out = torch.empty((11, n), dtype=torch.float64, device=dx.device)
    with torch.cuda.device(dx.device):
        rc = lib.dynpy_cart_to_dipolar_f64(dx.data_ptr(), dy.data_ptr(), dz.data_ptr(), dr.data_ptr(), n, out.data_ptr(),
                                           _stream())
    _lib.check(rc, "dynpy_cart_to_dipolar_f64")
    return out


def supercell(coords, frame: int, a: float):
    """[A,3,F] raw coords -> [3, 27A]: the 27 periodic images of frame index `frame` (neighbors._worker)."""
    lib = _lib.load()
    coords = _chk(coords, torch.float64, "coords")
    A, three, F = coords.shape
    out = torch.empty((3, 27 * A), dtype=torch.float64, device=coords.device)
    with torch.cuda.device(coords.device):
        rc = lib.dynpy_supercell_f64(coords.data_ptr(), A, F, int(frame), float(a), out.data_ptr(), _stream())
    _lib.check(rc, "dynpy_supercell_f64")
    return out


def topology_check(coords, rcov, bond_extra, expect, cell, dmax):
    """(mismatching pair-frames, first mismatching frame or -1) of the per-frame bond flags
    against `expect` [A, A] uint8 (upper triangle used)."""
    lib = _lib.load()
    coords = _chk(coords, torch.float64, "coords")
    rcov = _chk(rcov, torch.float64, "rcov")
    expect = _chk(expect, torch.uint8, "expect")
    A, three, F = coords.shape
    out = torch.zeros(2, dtype=torch.int64, device=coords.device)
    with torch.cuda.device(coords.device):
        rc = lib.dynpy_topology_check_f64(coords.data_ptr(), A, F, rcov.data_ptr(), float(bond_extra),
                                          expect.data_ptr(), float(cell[0]), float(cell[1]), float(cell[2]),
                                          float(dmax), out.data_ptr(), _stream())
    _lib.check(rc, "dynpy_topology_check_f64")
    bad, first = out.tolist()
    return int(bad), (int(first) if bad else -1)


def topology_events(coords, rcov, bond_extra, expect, cell, dmax, capacity=1 << 16, frames=None):
    """Pair-frames whose bond flag differs from `expect`: int64 array [k, 3] of (frame index, i, j), host.
    Raises when more than `capacity` differ (a trajectory that reactive is outside the static-topology design).
    frames = (begin, end): scan only that block of frames (one block per rank, the host merges the lists)."""
    lib = _lib.load()
    coords = _chk(coords, torch.float64, "coords")
    rcov = _chk(rcov, torch.float64, "rcov")
    expect = _chk(expect, torch.uint8, "expect")
    A, three, F = coords.shape
    out = torch.zeros(3, dtype=torch.int64, device=coords.device)
    ev = torch.empty((capacity, 3), dtype=torch.int64, device=coords.device)
    with torch.cuda.device(coords.device):
        f0, f1 = (0, F) if frames is None else (int(frames[0]), int(frames[1]))
        rc = lib.dynpy_topology_events_range_f64(coords.data_ptr(), A, F, f0, f1, rcov.data_ptr(), float(bond_extra),
                                                 expect.data_ptr(), float(cell[0]), float(cell[1]), float(cell[2]),
                                                 float(dmax), ev.data_ptr(), int(capacity), out.data_ptr(), _stream())
    _lib.check(rc, "dynpy_topology_events_range_f64")
    bad, first, n = out.tolist()
    if n > capacity:
        raise _lib.DynpyError("bond topology differs from the indexing frame in %d pair-frames (> %d): too reactive for "
                              "the static-topology indexing" % (n, capacity))
    if n == 0:
        return ev[:0].cpu().numpy()
    e = ev[:n]
    order = torch.argsort(e[:, 0] * (A * A) + e[:, 1] * A + e[:, 2])   # (frame, i, j) order, sorted on the device
    return e[order].cpu().numpy()


def dd_pair_count(coords, pair_i, pair_j, cell, rmax, frame_hit=None):
    """Frames in which each pair is inside rmax; optional int32 frame_hit[F] gets 1 where any pair is."""
    lib = _lib.load()
    coords = _chk(coords, torch.float64, "coords")
    pair_i = _chk(pair_i, torch.int32, "pair_i")
    pair_j = _chk(pair_j, torch.int32, "pair_j")
    A, three, F = coords.shape
    counts = torch.empty(pair_i.numel(), dtype=torch.int64, device=coords.device)
    with torch.cuda.device(coords.device):
        rc = lib.dynpy_dd_pair_count_f64(coords.data_ptr(), A, F, pair_i.data_ptr(), pair_j.data_ptr(), pair_i.numel(),
                                         float(cell[0]), float(cell[1]), float(cell[2]), float(rmax),
                                         counts.data_ptr(), frame_hit.data_ptr() if frame_hit is not None else None,
                                         _stream())
    _lib.check(rc, "dynpy_dd_pair_count_f64")
    return counts


def default_ws_budget(device) -> int:
    """Workspace of the fused dipolar call when the caller names none: a quarter of the free device memory,
    between 4 and 32 GiB.  Batches of a few items leave tails on 148 SMs: at 1e6 frames one item is 0.43 GB, and
    the step is 4 - 10 % faster with 16 - 32 GiB than with 4."""
    free, _ = torch.cuda.mem_get_info(device)
    return int(max(4 << 30, min(32 << 30, free // 4)))


def dd_pairs_to_spectrum(coords, pair_i, pair_j, cell, M=None, spectrum=None, ws_budget: int | None = None):
    """Fused dipolar path for always-present pairs: wrapped coords [A,3,F] -> spectrum [7, M]."""
    lib = _lib.load()
    coords = _chk(coords, torch.float64, "coords")
    pair_i = _chk(pair_i, torch.int32, "pair_i")
    pair_j = _chk(pair_j, torch.int32, "pair_j")
    A, three, F = coords.shape
    M = fft_length(F) if M is None else int(M)
    if spectrum is None:
        spectrum = torch.zeros((7, M), dtype=torch.float64, device=coords.device)
    _chk(spectrum, torch.float64, "spectrum")
    npairs = pair_i.numel()
    items = (npairs + 1) // 2
    per = int(lib.dynpy_dd_workspace_bytes(F, M, 1))
    if ws_budget is None:
        buf = _WS.get((coords.device.index, torch.cuda.current_stream(coords.device).cuda_stream))
        # an existing workspace counts as free memory: it is what this call is going to reuse
        ws_budget = max(default_ws_budget(coords.device), buf.numel() if buf is not None else 0)
    n = max(1, min(items, ws_budget // per))
    with torch.cuda.device(coords.device):
        wsb = int(lib.dynpy_dd_workspace_bytes(F, M, n))
        ws = workspace(coords.device, wsb)
        rc = lib.dynpy_dd_pairs_to_spectrum_f64(coords.data_ptr(), A, F, pair_i.data_ptr(), pair_j.data_ptr(), npairs,
                                                float(cell[0]), float(cell[1]), float(cell[2]), spectrum.data_ptr(), M,
                                                ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "dynpy_dd_pairs_to_spectrum_f64")
    return spectrum


def dd_pairs_ragged_to_acf(coords, pair_i, pair_j, cell, rmax, M=None, G=None, ws_budget: int = 8 << 30, mask=None,
                           pair_frames=None, frame_hit=None, counts=None):
    """Ragged dipolar path: G[7, F] += per-pair ACFs binned at the k-th present frame.  Optional mask [P, F]
    uint8 (frames in which the pair row passes the molecule filter), pair_frames [P] int64 out (frames present),
    frame_hit [F] int32 (set to 1 where some pair is present).
    counts [P] int64 (frames each pair is inside rmax, from dd_pair_count): the pairs are then processed by LENGTH
    CLASS — transform length and buffers sized by the class's largest count instead of F (not with mask /
    pair_frames, whose rows follow the caller's pair order)."""
    lib = _lib.load()
    coords = _chk(coords, torch.float64, "coords")
    pair_i = _chk(pair_i, torch.int32, "pair_i")
    pair_j = _chk(pair_j, torch.int32, "pair_j")
    A, three, F = coords.shape
    if G is None:
        G = torch.zeros((7, F), dtype=torch.float64, device=coords.device)
    _chk(G, torch.float64, "G")
    npairs = pair_i.numel()
    if mask is not None:
        mask = _chk(mask, torch.uint8, "mask")
        assert mask.shape == (npairs, F)
    if pair_frames is not None:
        pair_frames = _chk(pair_frames, torch.int64, "pair_frames")
    if frame_hit is not None:
        frame_hit = _chk(frame_hit, torch.int32, "frame_hit")

    def run(pi, pj, n_cap, Mc, msk, pf):
        n = pi.numel()
        one = int(lib.dynpy_dd_ragged_workspace_bytes(n_cap, Mc, 1))
        two = int(lib.dynpy_dd_ragged_workspace_bytes(n_cap, Mc, 2))
        per = two - one
        nb = max(1, min(n, (ws_budget - (one - per)) // per))
        with torch.cuda.device(coords.device):
            wsb = int(lib.dynpy_dd_ragged_workspace_bytes(n_cap, Mc, nb))
            while wsb > ws_budget and nb > 1:      # larger calls carry a larger T (not linear in the pair count)
                nb = max(1, nb - max(1, -(-(wsb - ws_budget) // per)))
                wsb = int(lib.dynpy_dd_ragged_workspace_bytes(n_cap, Mc, nb))
            ws = workspace(coords.device, wsb)
            rc = lib.dynpy_dd_pairs_ragged_capped_f64(coords.data_ptr(), A, F, pi.data_ptr(), pj.data_ptr(), n,
                                                      float(cell[0]), float(cell[1]), float(cell[2]), float(rmax),
                                                      msk.data_ptr() if msk is not None else None, G.data_ptr(),
                                                      pf.data_ptr() if pf is not None else None,
                                                      frame_hit.data_ptr() if frame_hit is not None else None,
                                                      int(n_cap), int(Mc), ws.data_ptr(), ws.numel(), _stream())
        _lib.check(rc, "dynpy_dd_pairs_ragged_capped_f64")

    import os
    if (counts is None or mask is not None or pair_frames is not None or npairs == 0
            or os.environ.get("DYNPY_RAGGED_CLASSES") == "0"):
        run(pair_i, pair_j, F, fft_length(F) if M is None else int(M), mask, pair_frames)
        return G
    # ---- length classes: pairs whose count needs the same transform length go together
    counts = _chk(counts, torch.int64, "counts")
    assert counts.numel() == npairs
    ch = counts.cpu().numpy()
    uniq, inv = np.unique(ch, return_inverse=True)
    mc = np.array([fft_length(int(c)) if c > 0 else 0 for c in uniq], dtype=np.int64)[inv]
    for Mc in sorted(set(mc.tolist()) - {0}, reverse=True):
        sel = torch.from_numpy(np.nonzero(mc == Mc)[0]).to(coords.device)
        n_cap = int(min(F, ch[mc == Mc].max()))
        run(pair_i[sel].contiguous(), pair_j[sel].contiguous(), n_cap, Mc, None, None)
    return G


# ------------------------------------------------------------------------------ SR
def sr_angmom(pos, mol_atoms, mass, axis_atoms, axis_rule, cell, inv_dt, vel=None, want_omega=True, want_axes=True):
    """pos [A,3,F] raw -> J [nmol,3,Fo] (+ omega [nmol,3,Fo], axes [nmol,9,Fo]); Fo = F-1 without vel."""
    lib = _lib.load()
    pos = _chk(pos, torch.float64, "pos")
    mol_atoms = _chk(mol_atoms, torch.int32, "mol_atoms")
    axis_atoms = _chk(axis_atoms, torch.int32, "axis_atoms")
    mass = _chk(mass, torch.float64, "mass")
    if vel is not None:
        vel = _chk(vel, torch.float64, "vel")
    A, three, F = pos.shape
    nmol, nat = mol_atoms.shape
    Fo = F if vel is not None else F - 1
    dev = pos.device
    J = torch.empty((nmol, 3, Fo), dtype=torch.float64, device=dev)
    om = torch.empty((nmol, 3, Fo), dtype=torch.float64, device=dev) if want_omega else None
    ax = torch.empty((nmol, 9, Fo), dtype=torch.float64, device=dev) if want_axes else None
    with torch.cuda.device(dev):
        rc = lib.dynpy_sr_angmom_f64(pos.data_ptr(), vel.data_ptr() if vel is not None else None, A, F,
                                     mol_atoms.data_ptr(), nmol, nat, mass.data_ptr(), axis_atoms.data_ptr(),
                                     int(axis_rule), float(cell[0]), float(cell[1]), float(cell[2]), float(inv_dt),
                                     J.data_ptr(), om.data_ptr() if om is not None else None,
                                     ax.data_ptr() if ax is not None else None, _stream())
    _lib.check(rc, "dynpy_sr_angmom_f64")
    return J, om, ax


# ------------------------------------------------------------------------------ Simpson
SIMPSON_MODES = {"cartwright": 0, "avg": 1}


def simpson(y, x, mode="cartwright", n=None):
    """scipy.integrate.simpson(y[c, :n], x=x[:n]) for every row c of y [cols, len] -> tensor [cols]."""
    lib = _lib.load()
    y = _chk(y, torch.float64, "y")
    x = _chk(x, torch.float64, "x")
    if y.dim() == 1:
        y = y[None]
    cols, stride = y.shape
    n = stride if n is None else int(n)
    if n < 1 or n > stride or x.numel() < n:
        raise _lib.DynpyError("simpson: n = %d outside the series (length %d)" % (n, stride))
    out = torch.empty(cols, dtype=torch.float64, device=y.device)
    with torch.cuda.device(y.device):
        rc = lib.dynpy_simpson_f64(y.data_ptr(), x.data_ptr(), n, cols, stride, SIMPSON_MODES[mode], out.data_ptr(),
                                   _stream())
    _lib.check(rc, "dynpy_simpson_f64")
    return out


# ------------------------------------------------------------------------------ J(omega)
def dft(re, im=None, n_out=None):
    """numpy.fft.fft along the last axis of re (+ i im) [S, n] for ANY n -> complex128 [S, n_out]."""
    lib = _lib.load()
    re = _chk(re, torch.float64, "re")
    if re.dim() == 1:
        re = re[None]
    S, n = re.shape
    if im is not None:
        im = _chk(im, torch.float64, "im")
        if im.dim() == 1:
            im = im[None]
        assert im.shape == re.shape
    n_out = n if n_out is None else int(n_out)
    out = torch.empty((S, n_out, 2), dtype=torch.float64, device=re.device)
    if out.numel() == 0:
        return torch.view_as_complex(out)
    with torch.cuda.device(re.device):
        wsb = int(lib.dynpy_dft_workspace_bytes(n))
        ws = workspace(re.device, wsb)
        rc = lib.dynpy_dft_f64(re.data_ptr(), im.data_ptr() if im is not None else None, S, n, out.data_ptr(), n_out,
                               ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "dynpy_dft_f64")
    return torch.view_as_complex(out)


def cutoff_bounds(y, window=100, tol=1e-3):
    """Per row of y [cols, n]: first index i < n-1 whose trailing `window`-mean of y/y[0] is within tol
    of zero, else n-1 -> int64 tensor [cols] (device)."""
    lib = _lib.load()
    y = _chk(y, torch.float64, "y")
    if y.dim() == 1:
        y = y[None]
    cols, n = y.shape
    out = torch.empty(cols, dtype=torch.int64, device=y.device)
    with torch.cuda.device(y.device):
        rc = lib.dynpy_cutoff_bounds_f64(y.data_ptr(), n, cols, n, int(window), float(tol), out.data_ptr(), _stream())
    _lib.check(rc, "dynpy_cutoff_bounds_f64")
    return out
