"""Quadrupolar driver — mirror of the reference's qrax.py (QR_module_main :28-96) on top of
the CUDA path: EFG CSV -> rank-2 spherical components -> Wiener–Khinchin ACFs summed over
nuclei -> Simpson J(0) -> 1/T1, 1/T2, 1/Tiso, tau_c, <V(0)^2>.  Same parameter class
(``Quadrupolar{data_set, analyte}``) and CSV outputs (``<stem>-<sym>-relax.csv``, ``-acfs.csv``)."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import dist, ops
from .config import default_simpson_mode
from .nuc import E_CHARGE, HBAR, nuc

F_COLUMNS = ['$f_{2,-2}$', '$f_{2,-1}$', '$f_{2,0}$', '$f_{2,1}$', '$f_{2,2}$']
_DTYPE = {'system': 'category', 'traj': 'i8', 'frame': 'i8', 'time': 'f8', 'label': 'i8', 'symbol': 'category',
          'Vxx': 'f8', 'Vxy': 'f8', 'Vxz': 'f8', 'Vyx': 'f8', 'Vyy': 'f8', 'Vyz': 'f8', 'Vzx': 'f8', 'Vzy': 'f8',
          'Vzz': 'f8', 'V11': 'f8', 'V22': 'f8', 'V33': 'f8', 'eta': 'f8'}
_COMPS = ['Vxx', 'Vyy', 'Vzz', 'Vxy', 'Vxz', 'Vyz']


def read_efg_data(file_path, dtype=None):
    """qrax.py:98-112"""
    return pd.read_csv(file_path, dtype=dtype)


def _interpolate(a):
    """pd.Series.interpolate() (linear, forward): interior and trailing NaNs are filled, leading ones stay."""
    ok = ~np.isnan(a)
    if ok.all() or not ok.any():
        return a
    idx = np.arange(len(a))
    first = idx[ok][0]
    a = a.copy()
    a[first:] = np.interp(idx[first:], idx[ok], a[ok])
    return a


def _fill_missing(v):
    """Missing EFG values of one nucleus, v [6, N] = Vxx, Vyy, Vzz, Vxy, Vxz, Vyz (input cleaning, host side;
    only runs when the series holds NaNs).  wiener_khinchin interpolates the REAL and IMAGINARY parts of each
    spherical component separately (qrax.py:149-150).  Which samples are missing follows the reference's complex
    arithmetic (qrax.py:115-130): Re R_2,+-2 = (Vxx - Vyy)/2 is missing where either term is, and `1j * NaN` is
    NaN in BOTH parts, so a missing Vxy (Vyz) also voids the real part of R_2,+-2 (R_2,+-1).  The five real
    series are interpolated and an equivalent Cartesian set (Vxx = -Vyy = Re R_22) goes to the device path."""
    if not np.isnan(v).any():
        return v
    out = np.empty_like(v)
    d = 0.5 * (v[0] - v[1])
    d[np.isnan(v[3])] = np.nan
    d = _interpolate(d)
    out[0], out[1] = d, -d
    xz = v[4].copy()
    xz[np.isnan(v[5])] = np.nan
    out[4] = _interpolate(xz)
    for c in (2, 3, 5):
        out[c] = _interpolate(v[c])
    return out


def wiener_khinchin(f):
    """qrax.py:147-156 (== ddrax.py:276-285, SRrax.py:216-225) for one series or a batch [S, N]: NaNs of the
    real and imaginary parts are interpolated separately (host, only when present), then the biased linear ACF
    real(ifft(|fft(f, 2N)|^2))[:N] / N comes from the CUDA Wiener–Khinchin kernels.  numpy in -> numpy out,
    CUDA tensor in -> CUDA tensor out (real or complex input)."""
    as_numpy = not isinstance(f, torch.Tensor)
    t = torch.as_tensor(np.asarray(f)) if as_numpy else f
    one = t.dim() == 1
    if one:
        t = t[None]
    re = (t.real if t.is_complex() else t).to(torch.float64)
    im = t.imag.to(torch.float64) if t.is_complex() else None
    if bool(torch.isnan(re).any()) or (im is not None and bool(torch.isnan(im).any())):
        re = torch.from_numpy(np.stack([_interpolate(r) for r in re.cpu().numpy()]))
        if im is not None:
            im = torch.from_numpy(np.stack([_interpolate(r) for r in im.cpu().numpy()]))
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    acf = ops.wk_acf_each(re.to(dev).contiguous(), None if im is None else im.to(dev).contiguous())
    acf = acf[0] if one else acf
    return acf.cpu().numpy() if as_numpy else acf


def efg_acf_mean(efg: torch.Tensor, shard=True):
    """efg [S,6,N] cuda -> mean over the S nuclei of the five ACFs f_{2,m}, m=-2..2: [5,N] cuda
    (cart_to_spatial + groupby('label').apply(correlate) + mean over labels, qrax.py:59-63)."""
    S, _, N = efg.shape
    M = ops.fft_length(N)
    lo, hi = dist.shard_range(S) if shard else (0, S)
    spec = torch.zeros((3, M), dtype=torch.float64, device=efg.device)
    if hi > lo:
        ops.qr_efg_to_spectrum(efg[lo:hi].contiguous(), M=M, spectrum=spec)
    if shard:
        dist.allreduce_sum_(spec)
    acf = ops.ifft_acf(spec, N, 1.0 / (float(M) * float(N) * float(S)))  # rows: R20, R2+-1, R2+-2
    return acf[[2, 1, 0, 1, 2]].contiguous()


def spectral_dens(time, f, simpson_mode="cartwright", dt=None, cutoff=False, cutoff_tol=1e-3):
    """qrax.py:169-201: g_2m = simpson(f_2m, x=time) (or dx=dt); g_iso = mean. Device tensors in.
    cutoff=True integrates up to the first lag whose trailing 100-point mean of f/f(0) is within cutoff_tol
    of zero — the bound of the LAST column (f_22) for all five, as the reference's leaked loop variable does
    (qrax.py:184)."""
    y = f.contiguous()
    if cutoff:
        b = int(ops.cutoff_bounds(y, 100, cutoff_tol)[-1])
        if b < 1:
            raise ValueError("cutoff bound %d leaves no samples to integrate" % b)
        g = ops.simpson(y, time.contiguous(), simpson_mode, n=b).cpu().numpy()
    else:
        x = time.contiguous()
        if dt:
            x = torch.arange(y.shape[1], dtype=torch.float64, device=y.device) * float(dt)
        g = ops.simpson(y, x, simpson_mode).cpu().numpy()
    return [float(x) for x in g], float(np.mean(g))


def relaxation(g, giso, analyte):
    """qrax.py:218-233"""
    quad_mom = nuc[analyte]['Q']
    s = nuc[analyte]['I']
    C_q = E_CHARGE * quad_mom / HBAR
    s_const = (2 * s + 3) / (s ** 2 * (2 * s - 1))
    au_q = 9.71736408e21
    gm2, gm1, g0_, gp1, gp2 = g
    g0 = 4 * gp2 + gp1 + gm1 + 4 * gm2
    g1 = 2 * gm2 + 3 * gm1 + 2 * gp1 + 3 * g0_
    gg = 10 * giso
    k = (1 / 40) * C_q ** 2 * s_const
    return k * g0 * au_q ** 2 * 1e-12, k * g1 * au_q ** 2 * 1e-12, k * gg * au_q ** 2 * 1e-12


def QR_module_main(QR, label=None, device=None):
    rawdf = read_efg_data(QR.data_set, dtype=_DTYPE)
    symbol = "".join([char for char in QR.analyte if char.isalpha()])
    simpson_mode = getattr(QR, 'simpson_mode', None) or default_simpson_mode()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    rank, world = dist.rank_world()
    ACFs, res = {}, {}
    for traj, df in rawdf.groupby('traj'):
        if traj == 0:
            continue
        df = df.sort_values(by='time')
        df['frame'] = df['frame'] - df.frame.iloc[0]
        df['time'] = df['time'] - df.time.iloc[0]
        if label:
            adf = df[df['label'] == int(label)]
        else:
            adf = df[df['symbol'] == symbol]
        labels = sorted(adf['label'].unique().tolist())
        series = {lab: adf[adf['label'] == lab] for lab in labels}
        lengths = {len(v) for v in series.values()}
        if len(lengths) != 1:
            raise ValueError("nuclei of trajectory %r have EFG series of different lengths %s" % (traj, sorted(lengths)))
        N = lengths.pop()
        host = np.empty((len(labels), 6, N), dtype=np.float64)
        for k, lab in enumerate(labels):
            host[k] = _fill_missing(np.stack([series[lab][name].to_numpy(dtype=np.float64) for name in _COMPS]))
        first = series[labels[0]]
        frame = np.mean([series[lab]['frame'].to_numpy(dtype=np.float64) for lab in labels], axis=0)
        tm = np.mean([series[lab]['time'].to_numpy(dtype=np.float64) for lab in labels], axis=0)
        f = efg_acf_mean(torch.from_numpy(host).to(device))
        tdev = torch.from_numpy(tm).to(device)
        g, giso = spectral_dens(tdev, f, simpson_mode)
        fh = f.cpu().numpy()
        v0 = float(fh[:, 0].sum())
        t1, t2, tiso = relaxation(g, giso, QR.analyte)
        acf_mean = pd.DataFrame({'frame': frame, 'time': tm})
        for c, name in enumerate(F_COLUMNS):
            acf_mean[name] = fh[c]
        acf_mean['symbol'] = symbol
        acf_mean.index = pd.Index(frame, name='frame')
        ACFs[traj] = acf_mean
        res[traj] = pd.DataFrame.from_dict({'symbol': [symbol], r'$\frac{1}{T_{1}}$': [t1],
                                            r'$\frac{1}{T_{2}}$': [t2], r'$\frac{1}{T_{iso}}$': [tiso],
                                            '$\\tau_{c}$': [giso * 5 / v0], r'$\langle V(0)^2\rangle$': [v0]})
        del first
    rax = pd.concat(res)
    rax.index = rax.index.droplevel(level=1)
    rax_mean = pd.DataFrame(rax.mean(numeric_only=True)).T
    rax_mean.index = ["mean"]
    rax_sem = pd.DataFrame(rax.sem(numeric_only=True)).T
    rax_sem.index = ["err"]
    all_rax = pd.concat([rax, rax_mean, rax_sem], sort=False)
    acfs = pd.concat(ACFs)
    system = QR.data_set.split('.csv')[0]
    if rank == 0:
        all_rax.to_csv(system + "-" + symbol + "-relax.csv", index_label="traj")
        acfs.to_csv(system + "-" + symbol + "-acfs.csv", index=False)
        print("Results written to " + system + "-" + symbol + "-relax.csv")
    return all_rax, acfs
