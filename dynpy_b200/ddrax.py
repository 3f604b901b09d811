"""Dipolar numerics — host-side mirror of the live part of the reference's ddrax.py.

``pair_acf_sum`` replaces pdist_ortho-per-frame + cart_to_dipolar_from_df +
groupby(['label0','label1']).apply(correlate) + the sum over pairs
(distances.py:131-145, ddrax.py:110-126,287-295, DDparse.py:81-104) with CUDA kernels;
``spec_dens_extr_narr`` and ``dipolar`` keep the reference's names and formulas
(ddrax.py:186-274)."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import dist, ops
from .nuc import HBAR, MU_0, nuc

G_COLUMNS = ['$G_{Y2,0}$', '$G_{Y2,1}$', '$G_{Y2,2}$', '$G_{r}$', '$G_{2,0}$', '$G_{2,1}$', '$G_{2,2}$']


def wiener_khinchin(f):
    """ddrax.py:276-285: the same function as qrax.wiener_khinchin."""
    from .qrax import wiener_khinchin as wk
    return wk(f)


DIPOLAR_COLUMNS = ["$r^{-3}$", "$Y_{2,0}$", "$Y_{2,1}$", "$Y_{2,2}$", "$F_{2,0}$", "$F_{2,1}$", "$F_{2,2}$"]


def cart_to_dipolar_from_df(atom_two):
    """ddrax.py:110-126 with the reference's signature: a table with columns dx, dy, dz, dr, time, label0, label1
    -> DataFrame(time, label0, label1, r^-3, Y_2m, F_2m) (Y_21, Y_22, F_21, F_22 complex).  The words come from
    the CUDA kernel (Cartesian closed forms of the angle formulas, <= 2e-13 relative)."""
    import pandas as pd
    dev = torch.device("cuda", torch.cuda.current_device())
    cols = [torch.tensor(np.asarray(atom_two[c], dtype=np.float64), device=dev)
            for c in ("dx", "dy", "dz", "dr")]
    w = ops.cart_to_dipolar(*cols).cpu().numpy()
    return pd.DataFrame.from_dict({"time": atom_two['time'], "label0": atom_two['label0'], "label1": atom_two['label1'],
                                   "$r^{-3}$": w[0], "$Y_{2,0}$": w[1], "$Y_{2,1}$": w[2] + 1j * w[3],
                                   "$Y_{2,2}$": w[4] + 1j * w[5], "$F_{2,0}$": w[6], "$F_{2,1}$": w[7] + 1j * w[8],
                                   "$F_{2,2}$": w[9] + 1j * w[10]})


def gapless(cell, rmax) -> bool:
    """True when every pair is inside rmax in every frame: min-image distance <= sqrt(a^2+b^2+c^2)/2."""
    half_diag2 = 0.25 * (cell[0] ** 2 + cell[1] ** 2 + cell[2] ** 2)
    return float(rmax) ** 2 > half_diag2 * (1.0 + 1e-9)


def pair_acf_sum(wrapped, pair_i, pair_j, cell, rmax, shard=True, masked=None):
    """Sum over pairs of the 7 per-pair ACFs with the reference's presence semantics.

    wrapped [A,3,F] wrapped coords (cuda), pair_i/pair_j int32 atom labels (i < j).
    Returns (G [7,F] float64 cuda — NOT yet scaled by 2/N_spins, hit [F] bool cuda,
    n_spins int, info dict).  With torch.distributed initialised and shard=True the pairs are
    split over the ranks and the result is all-reduced (every rank returns the total).
    masked = (mi, mj, mask [K,F] uint8): extra pairs that pass the reference's molecule filter only in the frames
    flagged by mask (pairs whose intra/inter status changes along the trajectory); handled by rank 0."""
    dev = wrapped.device
    A, _, F = wrapped.shape
    P = int(pair_i.numel())
    M = ops.fft_length(F)
    lo, hi = dist.shard_range(P) if shard else (0, P)
    pi, pj = pair_i[lo:hi].contiguous(), pair_j[lo:hi].contiguous()
    G = torch.zeros((7, F), dtype=torch.float64, device=dev)
    hit = torch.zeros(F, dtype=torch.int32, device=dev)
    label_used = torch.zeros(A, dtype=torch.int32, device=dev)
    info = dict(pairs=P, full=0, partial=0, absent=0, M=M)
    if hi > lo:
        if gapless(cell, rmax):
            full_i, full_j = pi, pj
            part_i = part_j = None
            hit.fill_(1)
            info["full"] = hi - lo
        else:
            counts = ops.dd_pair_count(wrapped, pi, pj, cell, rmax, frame_hit=hit)
            full = counts == F
            part = (counts > 0) & ~full
            full_i, full_j = pi[full].contiguous(), pj[full].contiguous()
            part_i, part_j = pi[part].contiguous(), pj[part].contiguous()
            part_counts = counts[part].contiguous()
            info["full"] = int(full.sum())
            info["partial"] = int(part.sum())
            info["absent"] = (hi - lo) - info["full"] - info["partial"]
        if full_i.numel():
            spec = ops.dd_pairs_to_spectrum(wrapped, full_i, full_j, cell, M=M)
            G += ops.ifft_acf(spec, F, 1.0 / (float(M) * float(F)))
            label_used[full_i.long()] = 1
            label_used[full_j.long()] = 1
        if part_i is not None and part_i.numel():
            ops.dd_pairs_ragged_to_acf(wrapped, part_i, part_j, cell, rmax, M=M, G=G, counts=part_counts)
            label_used[part_i.long()] = 1
            label_used[part_j.long()] = 1
    if masked is not None and masked[0].numel() and (not shard or dist.rank_world()[0] == 0):
        mi, mj, mk = masked
        nfr = torch.zeros(mi.numel(), dtype=torch.int64, device=dev)
        ops.dd_pairs_ragged_to_acf(wrapped, mi, mj, cell, rmax, M=M, G=G, mask=mk, pair_frames=nfr, frame_hit=hit)
        seen = nfr > 0
        label_used[mi[seen].long()] = 1
        label_used[mj[seen].long()] = 1
        info["partial"] += int(seen.sum())
        info["absent"] += int((~seen).sum())
        info["pairs"] += int(mi.numel())
    if shard:
        dist.allreduce_sum_(G)
        dist.allreduce_sum_(hit)
        dist.allreduce_sum_(label_used)
        cnt = torch.tensor([info["full"], info["partial"], info["absent"]], dtype=torch.int64, device=dev)
        dist.allreduce_sum_(cnt)
        info["full"], info["partial"], info["absent"] = (int(v) for v in cnt.tolist())
    n_spins = int((label_used > 0).sum())
    return G, hit > 0, n_spins, info


J_COLUMNS = [r'$\omega$', r'$log(\omega)$', r'$\omega^{\\frac{1}{2}}$', r'$J_{0}(\omega)$', r'$J_{1}(\omega)$',
             r'$J_{2}(\omega)$']


def fourier_T(f):
    """ddrax.py:169-173: np.fft.fft(f)[range(n//2)] for series of ANY length, on the device.
    f: real tensor [n] or [S, n] (cuda) -> complex128 [n//2] / [S, n//2]."""
    one = f.dim() == 1
    F = ops.dft(f.contiguous(), n_out=f.shape[-1] // 2)
    return F[0] if one else F


def fourier_freq(t):
    """ddrax.py:174-178: np.fft.fftfreq(n, dt)[range(n//2)] with dt = t[1] - t[0] (host array in, host array out;
    numpy forms k * (1/(n*dt)), so do we)."""
    t = np.asarray(t, dtype=np.float64)
    n = len(t)
    val = 1.0 / (n * (t[1] - t[0]))
    return np.arange(0, n // 2, dtype=int) * val


def spec_dens(time, G):
    """ddrax.py:178-184: finite-frequency spectral densities J_m(omega) = FFT(G_2m)[:n//2] on the frequency
    grid of the ACF.  time [n], G [7, n] device tensors (rows in G_COLUMNS order) -> DataFrame with the
    reference's columns (omega, log10 omega, sqrt omega, J_0, J_1, J_2; the J's complex)."""
    import pandas as pd
    freq = fourier_freq(time.cpu().numpy())
    Jw = fourier_T(G[4:7].contiguous()).cpu().numpy()
    with np.errstate(divide='ignore'):
        logf = np.log10(freq)
    return pd.DataFrame({J_COLUMNS[0]: freq, J_COLUMNS[1]: logf, J_COLUMNS[2]: np.sqrt(freq),
                         J_COLUMNS[3]: Jw[0], J_COLUMNS[4]: Jw[1], J_COLUMNS[5]: Jw[2]})


def spec_dens_extr_narr(time, G, simpson_mode="cartwright", dt=None, cutoff=False, cutoff_tol=1e-3):
    """ddrax.py:186-222: j_2m = 2*simpson(G_2m, x=time) (or dx=dt); <F^2_m> = G_2m(0); tau's.

    cutoff=True integrates only up to the first lag whose trailing 100-point mean of G/G(0) is within
    cutoff_tol of zero; as in the reference the bound found for the LAST column (G_22) is applied to all
    three columns (the loop variable `l` is read after the loop, ddrax.py:200).
    time [n] and G [7,n] are device tensors (rows in G_COLUMNS order); returns a dict of floats
    (plus 'bounds' when cutoff)."""
    y = G[4:7].contiguous()
    extra = {}
    if cutoff:
        bounds = ops.cutoff_bounds(y, 100, cutoff_tol).tolist()
        b = int(bounds[-1])
        if b < 1:
            raise ValueError("cutoff bound %d leaves no samples to integrate" % b)
        j = (2.0 * ops.simpson(y, time.contiguous(), simpson_mode, n=b)).cpu().numpy()
        extra['bounds'] = [int(v) for v in bounds]
    else:
        x = time.contiguous()
        if dt:
            x = torch.arange(y.shape[1], dtype=torch.float64, device=y.device) * float(dt)
        j = (2.0 * ops.simpson(y, x, simpson_mode)).cpu().numpy()
    f = G[4:7, 0].cpu().numpy()
    g = {'$j_{2,0}$': float(j[0]), '$j_{2,1}$': float(j[1]), '$j_{2,2}$': float(j[2])}
    g[r'$\langle F^{2}_{0}\rangle$'] = float(f[0])
    g[r'$\langle F^{2}_{1}\rangle$'] = float(f[1])
    g[r'$\langle F^{2}_{2}\rangle$'] = float(f[2])
    g[r'$\langle F(0)^{2}\rangle$'] = (float(f[1]) + float(f[2])) / 2
    g[r'$\tau_{0}$'] = g['$j_{2,0}$'] / float(f[0])
    g[r'$\tau_{1}$'] = g['$j_{2,1}$'] / float(f[1])
    g[r'$\tau_{2}$'] = g['$j_{2,2}$'] / float(f[2])
    g[r'$\tau_{c}$'] = (g[r'$\tau_{1}$'] + g[r'$\tau_{2}$']) / 2
    g.update(extra)
    return g


def finite_freq_j(J, larmor_freq=25):
    """ddrax.py:249-262: J_m at the Larmor frequency by a straight line in sqrt(omega) through rows 1 and 2 of
    the spec_dens table: j0 ~ J_0(w0), j02 ~ J_0(2 w0), j1 ~ J_1(w0), j2 ~ J_2(2 w0), w0 = larmor_freq * 1e6.
    Evaluated exactly as written there: intercept MINUS slope * sqrt(w0), with the signed fitted slope."""
    omega_0 = larmor_freq * 1e6
    sq = np.real(np.asarray(J[J_COLUMNS[2]]))

    def line(col):
        y = np.real(np.asarray(J[col]))
        m = (y[2] - y[1]) / (sq[2] - sq[1])
        return m, y[2] - m * sq[2]

    m0, b0 = line(J_COLUMNS[3])
    m1, b1 = line(J_COLUMNS[4])
    m2, b2 = line(J_COLUMNS[5])
    return dict(j0=b0 - m0 * np.sqrt(omega_0), j02=b0 - m0 * np.sqrt(2 * omega_0), j1=b1 - m1 * np.sqrt(omega_0),
                j2=b2 - m2 * np.sqrt(2 * omega_0))


def dipolar(spec_dens, symbol1, symbol2='H', extr_narr=True, single_J=False, larmor_freq=25, J=None):
    """ddrax.py:224-274.  extr_narr=True is the branch the CLI reaches.  The other two branches cannot run in
    the reference as written (single_J: `symbol` undefined at :237; extr_narr=False: `J`, `tc`, `f` undefined
    at :249/:274) — here single_J returns the one-row table the reference tries to build, indexed by symbol1,
    and extr_narr=False takes the spec_dens() table as `J` (NameError without it, like the reference) and
    returns (R, None, None)."""
    s = nuc[symbol1]['I']
    gamma1 = nuc[symbol1]['gamma']
    gamma2 = nuc[symbol2]['gamma']
    C_d = (MU_0 / (4 * math.pi)) ** 2 * (gamma1 ** 2) * (gamma2 ** 2) * (HBAR ** 2)
    s_const = s * (s + 1)
    au_m = 5.29177e-11
    tc = f = None
    if single_J:
        import pandas as pd
        j1 = j2 = spec_dens
        R = np.real(C_d * s_const * (j1 + (4 * j2)) * au_m ** (-6) * 1e-12)
        return pd.DataFrame({'$\\frac{1}{T_{1}}$': R}, index=[symbol1])
    elif extr_narr:
        j0 = j1 = j2 = float(np.mean([spec_dens['$j_{2,0}$'], spec_dens['$j_{2,1}$'], spec_dens['$j_{2,2}$']]))
        tc = spec_dens['$\\tau_{c}$']
        f = C_d * s_const * au_m ** (-6) * spec_dens[r'$\langle F^{2}_{0}\rangle$']
    else:
        if J is None:
            raise NameError("name 'J' is not defined (ddrax.py:249): pass the spec_dens() table as J=")
        ff = finite_freq_j(J, larmor_freq)
        j0, j1, j2 = ff['j0'], ff['j1'], ff['j2']
    if symbol1 == symbol2:
        R = float(np.real(C_d * s_const * (j1 + (4 * j2)) * au_m ** (-6) * 1e-12))
    else:
        R = (2 / 3) * C_d * s_const * 10 * float(np.real(j0)) * au_m ** (-6) * 1e-12
    return R, tc, f
