"""Nearest-molecule clusters around a source atom — mirror of the reference's
``neighbors.periodic_nearest_neighbors_by_atom`` / ``_create_super_universe`` / ``_worker``
(neighbors.py:22-199), the cluster extraction behind ``--inputs``.

Per frame the reference builds the 3x3x3 block of periodic images (a periodic cell of edge 3a),
runs ``pdist_ortho`` over all 27n atoms, detects bonds and molecules in the block, sorts the pair rows
that touch a source atom by distance and keeps the first hit per molecule.  Here the image block, the
wrap and the pair table come from the bit-exact CUDA kernels (``dynpy_supercell_f64``,
``dynpy_wrap_f64``, ``dynpy_pdist_ortho_f64`` — counted first, so the table is sized by its rows);
only the rows that are bonds or touch a source atom leave the device, and the bookkeeping on those few
hundred rows (union-find in networkx's discovery order, the sort, the first-hit-per-molecule filter)
is host code.  Every column of the reference's tables is reproduced, including the ``two`` row index
into the block's pair table and the molecule numbers.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch

from . import ops, topology
from .elements import SYM2COVRAD


def _source_atoms(source, A, symbols):
    """Super-atom indices of the source atoms: rows of the central image (prj == 13) selected by
    atom-table index (int), label list, or symbol (neighbors.py:82-90)."""
    base = 13 * A
    if isinstance(source, (int, np.int32, np.int64)):
        return [int(source)] if base <= int(source) < base + A else []
    if isinstance(source, (list, tuple)):
        want = set(int(s) for s in source)
        return [base + l for l in range(A) if l in want]
    return [base + l for l in range(A) if symbols[l] == source]


def periodic_nearest_neighbors_by_atom(traj, source, a, sizes, dmax=8.0, bond_extra=0.45, **radii):
    """traj: parseMD.Trajectory (raw coordinates [A,3,F] on the device, symbols, frame ids).
    Returns {'nearest': DataFrame(frame, idx, two, atom, molecule), nn: namespace(atom=DataFrame(symbol, x, y,
    z, frame)) for nn in sizes} with the reference's rows, order and values."""
    pos = traj.pos
    A, _, F = pos.shape
    dev = pos.device
    a = float(a)
    cell = a * 3
    n27 = 27 * A
    symbols27 = list(traj.symbols) * 27
    radmap = {s: SYM2COVRAD[s] for s in set(traj.symbols)}
    radmap.update(radii)
    rc = torch.tensor([radmap[s] for s in symbols27], dtype=torch.float64, device=dev)
    index = torch.arange(n27, dtype=torch.int64, device=dev)
    src = _source_atoms(source, A, traj.symbols)
    is_src = torch.zeros(n27, dtype=torch.bool, device=dev)
    if src:
        is_src[torch.tensor(src, dtype=torch.int64, device=dev)] = True
    sym_arr = np.array(symbols27, dtype=object)
    nearest_all = []
    clusters = {nn: [] for nn in sizes}
    for fi in range(F):
        fdx = int(traj.frames[fi])
        x27 = ops.supercell(pos, fi, a)                      # _worker: raw image coordinates
        u27 = ops.wrap(x27, cell)                            # unit-cell coordinates of the 3a cell
        _, _, _, dr, a0, a1, _ = ops.pdist_ortho(u27[0], u27[1], u27[2], cell, cell, cell, index, dmax)
        bonded = dr <= (rc[a0] + rc[a1] + bond_extra)        # _compute_bonds (distances.py:244-257)
        topo = topology.Topology(n27, a0[bonded].cpu().numpy(), a1[bonded].cpu().numpy(), symbols27)
        mol = topo.frame_molecule_id(topo.mol_rank)
        src_mols = list(dict.fromkeys(int(mol[s]) for s in src))      # .unique(): order of appearance
        touch = is_src[a0] | is_src[a1]
        two = torch.nonzero(touch).flatten()
        r_dr = dr[two].cpu().numpy()
        r_a0 = a0[two].cpu().numpy()
        r_a1 = a1[two].cpu().numpy()
        two = two.cpu().numpy()
        order = np.argsort(r_dr, kind="quicksort")           # DataFrame.sort_values("dr") default
        st_two = np.repeat(two[order], 2)                    # stack(): atom0 then atom1 of each row
        st_atom = np.stack([r_a0[order], r_a1[order]], axis=1).reshape(-1)
        keep = ~np.isin(st_atom, src)
        st_two, st_atom = st_two[keep], st_atom[keep]
        idx = np.arange(len(st_atom))
        st_mol = mol[st_atom]
        keep = ~np.isin(st_mol, src_mols)
        idx, st_two, st_atom, st_mol = idx[keep], st_two[keep], st_atom[keep], st_mol[keep]
        _, first = np.unique(st_mol, return_index=True)      # drop_duplicates('molecule', keep='first')
        first.sort()
        near = pd.DataFrame({"frame": fdx, "idx": idx[first], "two": st_two[first], "atom": st_atom[first],
                             "molecule": st_mol[first].astype(np.int64)})
        nearest_all.append(near)
        if len(near):                                       # the reference loops over nearest['frame'].unique()
            xyz = x27.cpu().numpy()
            for nn in sizes:
                mdxs = near["molecule"].tolist()[:nn]
                mdxs.append(src_mols[0])
                sel = np.isin(mol, mdxs)
                clusters[nn].append(pd.DataFrame({"symbol": sym_arr[sel], "x": xyz[0][sel], "y": xyz[1][sel],
                                                  "z": xyz[2][sel], "frame": fdx}))
    out = {"nearest": pd.concat(nearest_all, ignore_index=True)}
    for nn in sizes:
        out[nn] = SimpleNamespace(atom=pd.concat(clusters[nn], ignore_index=True))
    return out
