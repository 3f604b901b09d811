"""Seeded synthetic inputs of the BASELINE shapes.

Everything is closed-form in the frame index (sums of sinusoids with per-molecule
random parameters drawn once on the CPU), so a trajectory can be produced in
frame chunks, on any device, in any layout, and a chunk is identical no matter
how the chunking was done.  No oracle / reference code is involved here.

Layouts
-------
``traj``  : ``[A, 3, F]`` float64, bohr, *unwrapped* coordinates (the atom-major,
            frame-contiguous layout the CUDA kernels stream).
``efg``   : ``[S, 6, F]`` float64 with component order Vxx,Vyy,Vzz,Vxy,Vxz,Vyz.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

BOHR_PER_ANG = 1.0 / 0.529177  # the reference's conversion (parseMD.py:197-199)

# rigid-body geometries in Angstrom, first atom first (order matters: the
# reference labels atoms by their position inside a frame)
_WATER_BODY = np.array([[0.0, 0.0, 0.0],
                        [0.7569503, 0.5858823, 0.0],
                        [-0.7569503, 0.5858823, 0.0]])
_WATER_SYM = ["O", "H", "H"]
_t = 1.09 / math.sqrt(3.0)
_METHANE_BODY = np.array([[0.0, 0.0, 0.0],
                          [_t, _t, _t],
                          [_t, -_t, -_t],
                          [-_t, _t, -_t],
                          [-_t, -_t, _t]])
_METHANE_SYM = ["C", "H", "H", "H", "H"]

# CH3-C#N along z, nitrile carbon first (so that the methyl rows (R, C) = (0, 1) and (C, H1) = (1, 2) exist in
# the i < j pair table): C-C 1.458, C#N 1.157, C-H 1.09 Angstrom, H-C-C 109.5 degrees
_hz, _hr = 1.458 + 1.09 * math.cos(math.radians(70.5)), 1.09 * math.sin(math.radians(70.5))
_ACN_BODY = np.array([[0.0, 0.0, 0.0],
                      [0.0, 0.0, 1.458],
                      [_hr, 0.0, _hz],
                      [-0.5 * _hr, 0.5 * math.sqrt(3.0) * _hr, _hz],
                      [-0.5 * _hr, -0.5 * math.sqrt(3.0) * _hr, _hz],
                      [0.0, 0.0, -1.157]])
_ACN_SYM = ["C", "C", "H", "H", "H", "N"]

SPECIES = {
    "water": (_WATER_BODY, _WATER_SYM),
    "methane": (_METHANE_BODY, _METHANE_SYM),
    "acetonitrile": (_ACN_BODY, _ACN_SYM),
}


def water_celldm(nmol: int) -> float:
    """Cubic cell edge in bohr at the density of the reference's 64-water example
    (notebooks/dynpy_params_water_dipolar.py:11: 16.22 A... for 128 atoms / 3)."""
    return 16.22 * (nmol / 128.0) ** (1.0 / 3.0) * BOHR_PER_ANG


def methane_celldm(nmol: int) -> float:
    """Cubic cell edge in bohr at the density of notebooks/dynpy_params_methane.py:11
    (216.86 A for 200 CH4)."""
    return 216.86 * (nmol / 200.0) ** (1.0 / 3.0) * BOHR_PER_ANG


class RigidBox:
    """A box of rigid molecules that translate and tumble smoothly.

    centre_m(t) = lattice_m + drift*t + sum_k A_mk sin(w_mk t + p_mk)
    orient_m(t) = Rodrigues(sum_k B_mk sin(v_mk t + q_mk)) @ R0_m
    """

    def __init__(self, species: str, nmol: int, celldm: float, seed: int,
                 trans_amp: float = 0.35, rot_amp: float = 2.5, drift: float = 0.0,
                 nmodes: int = 3, wmin: float = 0.002, wmax: float = 0.05):
        body, sym = SPECIES[species]
        self.species = species
        self.nmol = nmol
        self.celldm = float(celldm)
        self.body = torch.tensor(body * BOHR_PER_ANG, dtype=torch.float64)
        self.nat_mol = len(sym)
        self.symbols = sym * nmol
        self.nat = self.nat_mol * nmol
        rng = np.random.default_rng(seed)
        ncell = int(math.ceil(nmol ** (1.0 / 3.0) - 1e-9))
        grid = np.stack(np.meshgrid(*(np.arange(ncell),) * 3, indexing="ij"), -1).reshape(-1, 3)
        grid = grid[rng.permutation(len(grid))[:nmol]]
        spacing = celldm / ncell
        self.lattice = torch.tensor((grid + 0.5) * spacing + rng.uniform(-0.05, 0.05, (nmol, 3)) * spacing)
        self.A = torch.tensor(rng.uniform(-1, 1, (nmol, nmodes, 3)) * trans_amp)
        self.w = torch.tensor(rng.uniform(wmin, wmax, (nmol, nmodes)) * 2 * math.pi)
        self.p = torch.tensor(rng.uniform(0, 2 * math.pi, (nmol, nmodes)))
        self.B = torch.tensor(rng.uniform(-1, 1, (nmol, nmodes, 3)) * rot_amp)
        self.v = torch.tensor(rng.uniform(wmin, wmax, (nmol, nmodes)) * 2 * math.pi)
        self.q = torch.tensor(rng.uniform(0, 2 * math.pi, (nmol, nmodes)))
        q0 = rng.normal(size=(nmol, 4))
        q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
        self.R0 = torch.tensor(_quat_to_mat(q0))
        self.drift = torch.tensor(np.asarray([1.0, -0.7, 0.4]) * drift)

    def to(self, device):
        for k in ("body", "lattice", "A", "w", "p", "B", "v", "q", "R0", "drift"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    def frames(self, f0: int, f1: int, atoms: slice | None = None) -> torch.Tensor:
        """Coordinates of frames [f0, f1) as ``[A, 3, f1-f0]`` (bohr, unwrapped)."""
        dev = self.lattice.device
        t = torch.arange(f0, f1, dtype=torch.float64, device=dev)              # [T]
        ph = self.w[:, :, None] * t + self.p[:, :, None]                          # [M,K,T]
        cen = self.lattice[:, :, None] + self.drift[None, :, None] * t \
            + torch.einsum("mkc,mkt->mct", self.A, torch.sin(ph))                # [M,3,T]
        rv = torch.einsum("mkc,mkt->mct", self.B,
                          torch.sin(self.v[:, :, None] * t + self.q[:, :, None]))  # [M,3,T]
        ang = torch.sqrt((rv * rv).sum(1) + 1e-300)                               # [M,T]
        ax = rv / ang[:, None, :]
        c, s = torch.cos(ang), torch.sin(ang)
        b0 = torch.einsum("mij,aj->mai", self.R0, self.body)                      # [M,a,3]
        # Rodrigues: b cos + (k x b) sin + k (k.b)(1-cos), for every atom of the body
        kb = torch.einsum("mct,mac->mat", ax, b0)                                 # [M,a,T]
        kx = torch.stack([ax[:, 1, None, :] * b0[:, :, 2, None] - ax[:, 2, None, :] * b0[:, :, 1, None],
                          ax[:, 2, None, :] * b0[:, :, 0, None] - ax[:, 0, None, :] * b0[:, :, 2, None],
                          ax[:, 0, None, :] * b0[:, :, 1, None] - ax[:, 1, None, :] * b0[:, :, 0, None]], 2)
        rot = b0[:, :, :, None] * c[:, None, None, :] + kx * s[:, None, None, :] \
            + ax[:, None, :, :] * (kb * (1 - c)[:, None, :])[:, :, None, :]       # [M,a,3,T]
        out = (cen[:, None, :, :] + rot).reshape(self.nat, 3, f1 - f0)
        return out if atoms is None else out[atoms]

    def trajectory(self, nframes: int, chunk: int = 8192, atom_mask=None, out=None) -> torch.Tensor:
        """Whole trajectory ``[A(sel), 3, F]``; ``atom_mask`` keeps only some atoms resident."""
        dev = self.lattice.device
        idx = None if atom_mask is None else torch.as_tensor(atom_mask, device=dev)
        na = self.nat if idx is None else int(idx.numel() if idx.dtype != torch.bool else idx.sum())
        if out is None:
            out = torch.empty((na, 3, nframes), dtype=torch.float64, device=dev)
        for f0 in range(0, nframes, chunk):
            f1 = min(nframes, f0 + chunk)
            blk = self.frames(f0, f1)
            out[:, :, f0:f1] = blk if idx is None else blk[idx]
        return out


def _quat_to_mat(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)


def water_box(nmol, seed=2, **kw) -> RigidBox:
    return RigidBox("water", nmol, water_celldm(nmol), seed, **kw)


def methane_box(nmol, seed=3, **kw) -> RigidBox:
    kw.setdefault("trans_amp", 1.0)
    return RigidBox("methane", nmol, methane_celldm(nmol), seed, **kw)


def acetonitrile_box(nmol, seed=4, spacing=11.0, **kw) -> RigidBox:
    """Dilute box (lattice spacing in bohr) so that no intermolecular contact comes near a bond cut-off."""
    ncell = int(math.ceil(nmol ** (1.0 / 3.0) - 1e-9))
    kw.setdefault("trans_amp", 0.3)
    return RigidBox("acetonitrile", nmol, spacing * ncell, seed, **kw)


def efg_series(nnuc: int, nframes: int, seed: int = 1, device="cpu", nmodes: int = 6,
               f0: int = 0) -> torch.Tensor:
    """Symmetric traceless EFG tensors ``[S, 6, F]`` (Vxx,Vyy,Vzz,Vxy,Vxz,Vyz):
    five independent components, each a sum of random sinusoids (smooth memory)."""
    rng = np.random.default_rng(seed)
    a = torch.tensor(rng.normal(size=(nnuc, 5, nmodes)) / math.sqrt(nmodes), device=device)
    w = torch.tensor(rng.uniform(0.001, 0.08, (nnuc, 5, nmodes)) * 2 * math.pi, device=device)
    p = torch.tensor(rng.uniform(0, 2 * math.pi, (nnuc, 5, nmodes)), device=device)
    t = torch.arange(f0, f0 + nframes, dtype=torch.float64, device=device)
    out = torch.empty((nnuc, 6, nframes), dtype=torch.float64, device=device)
    for c in range(5):  # comps 0:Vxx 1:Vyy 3:Vxy 4:Vxz 5:Vyz ; Vzz = -(Vxx+Vyy)
        dst = (0, 1, 3, 4, 5)[c]
        out[:, dst, :] = torch.einsum("sk,skt->st", a[:, c], torch.sin(w[:, c, :, None] * t + p[:, c, :, None]))
    out[:, 2, :] = -(out[:, 0, :] + out[:, 1, :])
    return out


# ------------------------------------------------------------------ file writers
def write_xyz(path: str, traj: torch.Tensor, symbols, first_printed: int = 1) -> None:
    """Write ``[A,3,F]`` bohr coordinates as an .xyz trajectory in Angstrom
    (what parseMD.parse_xyz reads: nat line, comment line, ``sym x y z`` rows)."""
    x = (traj.detach().cpu().numpy() / BOHR_PER_ANG).transpose(2, 0, 1)  # [F,A,3]
    nat = x.shape[1]
    with open(path, "w") as fh:
        for f in range(x.shape[0]):
            fh.write(f"{nat}\nframe {f + first_printed}\n")
            for a in range(nat):
                fh.write(f"{symbols[a]} {x[f, a, 0]:.10f} {x[f, a, 1]:.10f} {x[f, a, 2]:.10f}\n")


def write_efg_csv(path: str, efg: torch.Tensor, symbol: str = "I", frame_step: int = 200,
                  dt: float = 0.029 / 200, ntraj: int = 1, system: str = "synthetic") -> None:
    """EFG CSV in the schema qrax.read_efg_data expects (qrax.py:29); ``efg`` is
    ``[S,6,F]`` per trajectory (the same series is reused, sign-flipped, for traj>1)."""
    e = efg.detach().cpu().numpy()
    S, _, F = e.shape
    cols = "system,traj,frame,time,label,symbol,Vxx,Vxy,Vxz,Vyx,Vyy,Vyz,Vzx,Vzy,Vzz\n"
    with open(path, "w") as fh:
        fh.write(cols)
        for tr in range(1, ntraj + 1):
            sgn = 1.0 if tr % 2 else -0.9
            for f in range(F):
                fr = 1000 + f * frame_step
                for s in range(S):
                    xx, yy, zz, xy, xz, yz = (sgn * e[s, :, f]).tolist()
                    fh.write(f"{system},{tr},{fr},{fr * dt!r},{s},{symbol},{xx!r},{xy!r},{xz!r},"
                             f"{xy!r},{yy!r},{yz!r},{xz!r},{yz!r},{zz!r}\n")


def write_params(path: str, body: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(body)
