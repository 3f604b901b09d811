"""Spin-rotation driver — mirror of the reference's SRparse.py (SR_module_main :25-106,
prep_SR_uni1/2/3 :109-238, compute_SR_rax :240-262) on top of the CUDA path.  Same parameter
classes and CSV outputs (``<traj_dir><traj>/{molax,ang_vel_molax,J_cart,Jacfs_all,Jacf}.csv``,
``<traj_dir>SR-results.csv``)."""
from __future__ import annotations

import sys
import time

import numpy as np
import pandas as pd
import torch

from . import dist, ops, parseMD, topology
from .config import default_simpson_mode
from .nuc import HBAR

J_COLUMNS = ['$J_{x}$', '$J_{y}$', '$J_{z}$']
_FORMULA = {'water': 'H(2)O(1)', 'acetonitrile': 'H(3)C(2)N(1)', 'methane': 'H(4)C(1)'}


def frame_step(frames: np.ndarray, nat: int):
    """u.atom.frame.diff().unique()[-1] (SRparse.py:132): last distinct value of the row-wise diff."""
    uniq = []
    for d in ([0] if nat > 1 else []) + np.diff(frames).tolist():
        if d not in uniq:
            uniq.append(d)
    return uniq[-1]


def select_molecules(ft, SR):
    """Molecules the reference keeps (SRparse.py:142-176): exact formula match in a frame, labelled by the rank of
    their atom-label tuple among all such tuples of all frames, present (as exactly that connected component)
    in EVERY frame, label < nmol.  `ft`: topology.FrameTopologies.  Returns list of (molecule_label, mol_rank in
    the indexing frame, atoms) in molecule order."""
    mol_type = SR.mol_type
    if mol_type in _FORMULA:
        ident = _FORMULA[mol_type]
    elif mol_type in ('methyl', 'ring'):
        ident = SR.identifier
    else:
        sys.exit("Only molecule types methane, water, acetonitrile, and methyl are supported")
    topo = ft.base
    hit = topo.classify_exact(ident)
    base = {topo.groups[m]: m for m in range(topo.nmol) if hit[m]}
    every = set(base)
    tuples = set(base)
    for t in ft.frames.values():
        try:
            h = t.classify_exact(ident)
        except KeyError:
            h = np.zeros(t.nmol, dtype=bool)
        here = {t.groups[m] for m in range(t.nmol) if h[m]}
        tuples |= here
        every &= here
    label_of = {t: k for k, t in enumerate(sorted(tuples))}
    keep = [(label_of[t], base[t], t) for t in every if label_of[t] < SR.nmol]
    keep.sort(key=lambda k: k[1])  # groupby('molecule') order: by molecule id inside the frame
    return keep, topology.atoms_per_formula(ident)


def axis_atoms_for(mol_type, atoms, two0, methyl_indeces=None):
    """(a0,b0,a1,b1): the two intramolecular pair rows the axis rule reads (SRrax.py:371-436)."""
    if mol_type == 'water':  # rows with mol-atom_index (0,1) and (0,2)
        return (atoms[0], atoms[1], atoms[0], atoms[2]), 1
    if mol_type == 'methane':  # first two intramolecular rows of atom_two in row-major i<j order
        s = set(atoms)
        rows = [(int(i), int(j)) for i, j in zip(two0["i"], two0["j"]) if int(i) in s and int(j) in s][:2]
        return (rows[0][0], rows[0][1], rows[1][0], rows[1][1]), 0
    if mol_type == 'methyl':
        R, C, H1 = methyl_indeces[0], methyl_indeces[1], methyl_indeces[2]
        return (atoms[R], atoms[C], atoms[C], atoms[H1]), 0
    if mol_type == 'ring':
        R, C = methyl_indeces[0], methyl_indeces[1]
        return (atoms[R], atoms[C], atoms[R], atoms[1]), 0
    sys.exit("Only molecule types methane, water, acetonitrile, and methyl are supported")


def compute_SR_rax(time_dev, acf_mean, SR, simpson_mode):
    """SRparse.py:240-262"""
    C = [float(c) * 1000 * 2 * np.pi for c in SR.C_SR]
    G = ops.simpson(acf_mean.contiguous(), time_dev.contiguous(), simpson_mode).cpu().numpy()
    a0 = acf_mean[:, 0].cpu().numpy()
    tx, ty, tz = (float(G[c] / a0[c]) for c in range(3))
    r = 2 / 3 / (HBAR ** 2) * (G[0] * C[0] ** 2 + G[1] * C[1] ** 2 + G[2] * C[2] ** 2) * (1e12) * (5.29177e-11) ** 4 \
        * (1.66054e-27) ** 2
    return tx, ty, tz, float(r)


def sr_trajectory(tr, PD, SR, simpson_mode, per_molecule_acfs=True):
    """One trajectory -> dict of result tables (host) and rates."""
    cell = (float(PD.celldm),) * 3
    dev = tr.pos.device
    methyl_indeces = getattr(SR, 'methyl_indeces', None)
    global_momentum = getattr(SR, 'global_momentum', False)
    if tr.vel is None:
        print("Explicit velocities not provided. Will be determined from position and timestep. If timestep is "
              "too large, these velocities will be inaccurate.")
        inv_dt = 1.0 / (frame_step(tr.frames, tr.nat) * PD.timestep)
        frames = tr.frames[1:]
        first = 1
    else:
        inv_dt = 0.0
        frames = tr.frames
        first = 0
    F = len(frames)
    wrapped = ops.wrap(tr.pos[:, :, first:].contiguous(), cell[0])
    ft, two0 = topology.build_topology(wrapped, tr.symbols, cell, 8.0, bond_extra=0.45, dynamic=True)
    topo = ft.base
    keep, nat_per_mol = select_molecules(ft, SR)
    if not keep:
        raise ValueError("no molecule of type %r found" % SR.mol_type)
    nmol = len(keep)
    mol_atoms, axis_atoms, rule = [], [], 0
    for lab, m, atoms in keep:
        use = list(atoms)
        ax4, rule = axis_atoms_for(SR.mol_type, atoms, two0, methyl_indeces)
        if SR.mol_type == 'methyl' and not global_momentum:
            use = [atoms[k] for k in sorted(methyl_indeces)]
        mol_atoms.append(use)
        axis_atoms.append(ax4)
    # one per-slot mass vector serves every molecule: the element order inside each kept molecule (ascending
    # atom label) must therefore be the same; the reference maps the mass per atom row (SRrax.py:448-451)
    sym0 = [tr.symbols[a] for a in mol_atoms[0]]
    for use in mol_atoms[1:]:
        if [tr.symbols[a] for a in use] != sym0:
            raise ValueError("molecules of type %r list their elements in different orders (%s vs %s): "
                             "per-slot masses would be wrong" % (SR.mol_type, sym0, [tr.symbols[a] for a in use]))
    mass = torch.tensor(topo.masses(mol_atoms[0]), dtype=torch.float64, device=dev)
    J, om, ax = ops.sr_angmom(tr.pos, torch.tensor(mol_atoms, dtype=torch.int32, device=dev), mass,
                              torch.tensor(axis_atoms, dtype=torch.int32, device=dev), rule, cell, inv_dt, vel=tr.vel)
    # ---- ACFs: mean over molecules (accumulated spectrum) and, for Jacfs_all.csv, per molecule
    lo, hi = dist.shard_range(nmol)
    M = ops.fft_length(F)
    spec = torch.zeros((3, M), dtype=torch.float64, device=dev)
    for c in range(3):
        if hi > lo:
            ops.wk_acf_sum(J[lo:hi, c, :], pack_real_pairs=True, M=M, spectrum=spec[c])  # strided view, no copy
    dist.allreduce_sum_(spec)
    acf_mean = ops.ifft_acf(spec, F, 1.0 / (float(M) * float(F) * float(nmol)))
    tdev = torch.from_numpy(frames).to(dev).to(torch.float64) * float(PD.timestep)
    tx, ty, tz, r = compute_SR_rax(tdev, acf_mean, SR, simpson_mode)
    out = dict(frames=frames, keep=keep, J=J, omega=om, axes=ax, acf_mean=acf_mean, time=tdev, tau=(tx, ty, tz), rate=r,
               topo=topo, frame_topologies=ft, nframes_total=F)
    if per_molecule_acfs:
        out["acf_each"] = ops.wk_acf_each(J.reshape(nmol * 3, F).contiguous(), M=M).reshape(nmol, 3, F)
    return out


def _tables(out):
    """Host DataFrames with the reference's row order and columns."""
    frames, keep = out["frames"], out["keep"]
    F, nmol = len(frames), len(keep)
    mol_ids = out["frame_topologies"].molecule_ids([atoms[0] for (_, _, atoms) in keep])  # [F, nmol]
    labels = np.array([lab for (lab, _, _) in keep])
    fr_col = np.repeat(frames, nmol)
    mol_col = mol_ids.reshape(-1)
    lab_col = np.tile(labels, F)
    zero_idx = np.zeros(F * nmol, dtype=np.int64)

    def fm(t, k):  # [nmol, k, F] -> [F*nmol, k]
        return t.permute(2, 0, 1).reshape(F * nmol, k).cpu().numpy()

    molax = pd.DataFrame(fm(out["axes"], 9), columns=['xx', 'xy', 'xz', 'yx', 'yy', 'yz', 'zx', 'zy', 'zz'], index=zero_idx)
    av = pd.DataFrame(fm(out["omega"], 3), columns=['x', 'y', 'z'], index=zero_idx)
    Jc = pd.DataFrame(fm(out["J"], 3), columns=['x', 'y', 'z'], index=zero_idx)
    for df in (molax, av, Jc):
        df['frame'] = fr_col
        df['molecule'] = mol_col
        df['molecule_label'] = lab_col
    tabs = dict(molax=molax, ang_vel_molax=av, J_cart=Jc)
    if "acf_each" in out:
        order = np.argsort(labels, kind="stable")  # groupby('molecule_label') order
        a = out["acf_each"][torch.from_numpy(order).to(out["acf_each"].device)]  # [nmol,3,F]
        Ja = pd.DataFrame(a.permute(0, 2, 1).reshape(nmol * F, 3).cpu().numpy(), columns=J_COLUMNS,
                          index=np.zeros(nmol * F, dtype=np.int64))
        Ja['frame'] = np.tile(frames, nmol)
        Ja['molecule_label'] = np.repeat(labels[order], F)
        Ja['molecule'] = mol_ids[:, order].T.reshape(-1)
        tabs["Jacfs_all"] = Ja
    am = out["acf_mean"].cpu().numpy()
    Jm = pd.DataFrame({J_COLUMNS[0]: am[0], J_COLUMNS[1]: am[1], J_COLUMNS[2]: am[2]})
    Jm['frame'] = frames.astype(np.float64)
    Jm['molecule_label'] = float(np.mean(labels))
    Jm['molecule'] = mol_ids.mean(axis=1)
    Jm['time'] = Jm['frame'] * 1.0
    Jm.index = pd.Index(frames, name='frame')
    tabs["Jacf"] = Jm
    return tabs


def SR_module_main(PD, SR):
    outer_start_time = time.time()
    simpson_mode = getattr(SR, 'simpson_mode', None) or default_simpson_mode()
    rank, world = dist.rank_world()
    say = print if rank == 0 else (lambda *a, **k: None)   # user-facing lines once, not once per rank
    res = {}
    for t, traj in enumerate(PD.trajs):
        inner_start_time = time.time()
        tr = parseMD.PARSE_MD(PD, traj)
        out = sr_trajectory(tr, PD, SR, simpson_mode)
        tabs = _tables(out)
        tabs["Jacf"]['time'] = tabs["Jacf"]['frame'] * PD.timestep
        if rank == 0:
            for name in ("molax", "ang_vel_molax", "J_cart", "Jacfs_all", "Jacf"):
                tabs[name].to_csv(PD.traj_dir + traj + '/' + name + '.csv')
        tx, ty, tz = out["tau"]
        r = out["rate"]
        say("1/T1     =     {t:.4f} Hz".format(t=r))
        say("Total SR Run Time for {s}    --- {t:.2f} seconds ---".format(s=PD.traj_dir + traj,
                                                                            t=time.time() - inner_start_time))
        res[traj] = tx, ty, tz, r
    res_df = pd.DataFrame(res).T
    res_df.columns = ["tau_x", "tau_y", "tau_z", "1/T1"]
    if rank == 0:
        res_df.to_csv(PD.traj_dir + 'SR-results.csv')
    say("Total SR Run Time         --- {t:.2f} seconds ---".format(t=time.time() - outer_start_time))
    return res_df
