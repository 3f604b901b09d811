"""Bonds, molecules and labels — the indexing the reference derives with
compute_atom_two/_compute_bonds (distances.py:68-103,244-257), compute_molecule
(universe.py:411-451), Molecule.classify (universe.py:340-384) and get_atom_labels
(universe.py:107-118).

The reference rebuilds a networkx graph over the atoms of ALL frames.  Here the pair list of
a frame comes from the bit-exact CUDA pdist_ortho kernel, the (tiny) union-find runs on the
host once, and the bond list is re-checked on a few other frames: a molecular-dynamics
topology that changes mid-trajectory is reported instead of silently mis-indexed.
"""
from __future__ import annotations

import re

import numpy as np
import torch

from . import ops
from .elements import SYM2COVRAD, SYM2MASS


def parse_formula(identifier: str) -> dict:
    """'H(2)O(1)' -> {'H': 2, 'O': 1}"""
    out = {}
    for sym, cnt in re.findall(r"([A-Z][a-z]?)\((\d+)\)", identifier):
        out[sym] = out.get(sym, 0) + int(cnt)
    if not out:
        raise KeyError("cannot parse molecular formula %r" % identifier)
    return out


def atoms_per_formula(identifier: str) -> int:
    """nat_per_mol exactly as DDparse.py:151 / SRparse.py:158 compute it."""
    return int(np.sum([int(n) for n in identifier.replace('(', ')').split(')') if n.isnumeric()]))


def frame_pairs(wrapped: torch.Tensor, cell, frame: int, dmax: float):
    """pdist_ortho of one frame of wrapped coords [A,3,F] -> dict of host arrays
    (i, j = atom labels inside the frame, row-major upper triangle order)."""
    A, _, F = wrapped.shape
    idx = torch.arange(A, dtype=torch.int64, device=wrapped.device)
    dx, dy, dz, dr, a0, a1, prj = ops.pdist_ortho(wrapped[:, 0, frame], wrapped[:, 1, frame], wrapped[:, 2, frame],
                                                  cell[0], cell[1], cell[2], idx, dmax, stride=3 * F)
    return dict(dx=dx.cpu().numpy(), dy=dy.cpu().numpy(), dz=dz.cpu().numpy(), dr=dr.cpu().numpy(),
                i=a0.cpu().numpy(), j=a1.cpu().numpy(), prj=prj.cpu().numpy())


def bonded(two: dict, symbols, bond_extra: float = 0.45) -> np.ndarray:
    """bond flag per row: dr <= rcov(a) + rcov(b) + bond_extra (distances.py:244-257)."""
    rc = np.array([SYM2COVRAD[s] for s in symbols], dtype=np.float64)
    return two["dr"] <= (rc[two["i"]] + rc[two["j"]] + bond_extra)


class Topology:
    """Molecules of one frame in the reference's numbering.

    mol_rank[a] : rank of atom a's connected component among ALL components of the frame (single atoms
                  included), ordered by their smallest atom — networkx's connected_components order.
    n_isolated  : number of unbonded atoms.  compute_molecule (universe.py:426-437) first numbers those
                  0..n_isolated-1 and then overwrites them while walking connected_components (which yields
                  single-node components too), so the number the reference ends up with for a molecule of a
                  single frame is n_isolated + mol_rank (`frame_molecule_id`).
    groups      : list (by mol_rank) of atom-label tuples (ascending); built on first use.
    comp_min[a] : smallest atom of a's component (the canonical name of the component).
    """

    def __init__(self, nat: int, bi: np.ndarray, bj: np.ndarray, symbols):
        parent = list(range(nat))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        deg = np.zeros(nat, dtype=np.int64)
        for a, b in zip(bi.tolist(), bj.tolist()):
            deg[a] += 1
            deg[b] += 1
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)   # the root of a component is its smallest atom
        self.nat = nat
        self.symbols = list(symbols)
        self.deg = deg
        self.comp_min = np.array([find(a) for a in range(nat)], dtype=np.int64)
        self.bonds = set(zip(bi.tolist(), bj.tolist()))
        self._finish()

    def _finish(self):
        """mol_rank / nmol / n_isolated from comp_min and deg (vectorised: also the whole cost of a derived frame)."""
        is_min = self.comp_min == np.arange(self.nat)
        self.mol_rank = (np.cumsum(is_min) - 1)[self.comp_min]
        self.nmol = int(is_min.sum())
        self.n_isolated = int((self.deg == 0).sum())
        self._groups = None
        self._mol_bonds = None

    @property
    def groups(self):
        if self._groups is None:
            order = np.argsort(self.mol_rank, kind="stable")
            cuts = np.cumsum(np.bincount(self.mol_rank, minlength=self.nmol))[:-1]
            self._groups = [tuple(g.tolist()) for g in np.split(order, cuts)]
        return self._groups

    def bonds_of(self, m: int):
        """bonds (i, j) inside molecule m"""
        if self._mol_bonds is None:
            mb = {}
            for ij in self.bonds:
                mb.setdefault(int(self.mol_rank[ij[0]]), []).append(ij)
            self._mol_bonds = mb
        return self._mol_bonds.get(int(m), [])

    def derive(self, pairs) -> "Topology":
        """Topology of a frame whose bond list is this one's with the flags of `pairs` toggled: only the molecules
        that hold an end of a toggled pair are re-indexed (a handful of atoms), the renumbering of everything behind
        them is three vectorised passes."""
        mols = sorted({int(self.mol_rank[a]) for ij in pairs for a in ij})
        atoms = [a for m in mols for a in self.groups[m]]
        bonds = set()
        for m in mols:
            bonds.update(self.bonds_of(m))
        toggled = [(int(i), int(j)) for i, j in pairs]
        bonds.symmetric_difference_update(toggled)
        parent = {a: a for a in atoms}

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for a, b in bonds:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        t = Topology.__new__(Topology)
        t.nat, t.symbols = self.nat, self.symbols
        t.comp_min = self.comp_min.copy()
        for a in atoms:
            t.comp_min[a] = find(a)
        t.deg = self.deg.copy()
        for ij in toggled:
            d = -1 if ij in self.bonds else 1
            t.deg[ij[0]] += d
            t.deg[ij[1]] += d
        t.bonds = None          # full bond set of a derived frame: base.bonds ^ toggled (not needed by any caller)
        t._finish()
        t._parent, t._touched = self, atoms
        return t

    def frame_molecule_id(self, m):
        """Number compute_molecule gives molecule(s) m when the universe holds this one frame."""
        return self.n_isolated + m

    def formula(self, m: int) -> dict:
        out = {}
        for a in self.groups[m]:
            out[self.symbols[a]] = out.get(self.symbols[a], 0) + 1
        return out

    def classify_exact(self, identifier) -> np.ndarray:
        """bool per molecule: formula == identifier (Molecule.classify with exact=True).
        Raises KeyError like the reference when nothing matches (universe.py:378)."""
        want = identifier if isinstance(identifier, dict) else parse_formula(identifier)
        syms = sorted(set(self.symbols) | set(want))
        code = {s: k for k, s in enumerate(syms)}
        cnt = np.zeros((self.nmol, len(syms)), dtype=np.int64)
        np.add.at(cnt, (self.mol_rank, np.array([code[s] for s in self.symbols])), 1)
        hit = (cnt == np.array([want.get(s, 0) for s in syms])[None, :]).all(axis=1)
        if not hit.any():
            raise KeyError('No records found for {}, with identifier {}.'.format("analyte", want))
        return hit

    def molecule_id(self, frame_index: int, m: int, nframes: int) -> int:
        """Global 'molecule' id the reference assigns (universe.py:426-437) in a universe of `nframes` frames
        with this topology in every frame: numbering starts after all isolated atoms of all frames, then every
        component (single atoms included) in atom order, frame by frame."""
        return nframes * self.n_isolated + frame_index * self.nmol + m

    def masses(self, atoms) -> np.ndarray:
        return np.array([SYM2MASS[self.symbols[a]] for a in atoms], dtype=np.float64)


class FrameTopologies:
    """The topology of the indexing frame plus the (few) frames whose bond list differs from it — what the
    reference's per-frame compute_molecule sees (universe.py:411-451).

    changes          {frame index: [(i, j), ...]} bond flags that differ from the indexing frame's
    frames           {frame index: Topology} of the deviating frames — derived from `base` on first use (a run that
                     never filters by molecule, pair_type 'all', does not pay for them)
    at(f)            Topology of frame index f
    molecule_ids(..) the 'molecule' numbers of the reference's all-frames universe: numbering starts after the
                     isolated atoms of ALL frames, then the components of frame 0, frame 1, ... in atom order
    """

    def __init__(self, base: Topology, changes: dict, nframes: int):
        self.base = base
        self.nframes = nframes
        self.changes = changes
        self._frames = None
        self._start_ = None

    @property
    def frames(self) -> dict:
        if self._frames is None:
            self._frames = {f: self.base.derive(pairs) for f, pairs in self.changes.items()}
        return self._frames

    @property
    def _start(self):
        if self._start_ is None:
            ncomp = np.full(self.nframes, self.base.nmol, dtype=np.int64)
            niso = np.full(self.nframes, self.base.n_isolated, dtype=np.int64)
            for f, t in self.frames.items():
                ncomp[f], niso[f] = t.nmol, t.n_isolated
            self._start_ = int(niso.sum()) + np.concatenate([[0], np.cumsum(ncomp)[:-1]])
        return self._start_

    def at(self, f: int) -> Topology:
        return self.frames[f] if f in self.changes else self.base

    def molecule_ids(self, first_atoms) -> np.ndarray:
        """[F, len(first_atoms)]: id, in every frame, of the molecule that holds each of the given atoms."""
        first_atoms = np.asarray(first_atoms, dtype=np.int64)
        out = self._start[:, None] + self.base.mol_rank[first_atoms][None, :]
        for f, t in self.frames.items():
            out[f] = self._start[f] + t.mol_rank[first_atoms]
        return out


class _LazyPairs(dict):
    """{frame: [(i, j), ...]} over an event array sorted by frame: the per-frame lists are cut out on first access
    (a reactive synthetic box has ~1e6 events; a run that never looks at them — pair_type 'all' — pays nothing)."""

    def __init__(self, ev: np.ndarray):
        super().__init__()
        self._ev = ev
        if len(ev):   # rows are sorted by frame: group boundaries in one pass, no sort
            f = ev[:, 0]
            self._first = np.flatnonzero(np.concatenate([[True], f[1:] != f[:-1]]))
            self._fr = f[self._first]
        else:
            self._first = np.empty(0, np.int64)
            self._fr = np.empty(0, np.int64)
        self._span_ = None

    @property
    def _span(self):
        """{frame: (begin, end)} — a python dict over ~1e5 frames costs ~15 ms, so it is only built when a caller
        looks frames up or iterates (len() and bool() do not need it)"""
        if self._span_ is None:
            last = np.append(self._first[1:], len(self._ev))
            self._span_ = dict(zip(self._fr.tolist(), zip(self._first.tolist(), last.tolist())))
        return self._span_

    def __len__(self):
        return int(len(self._fr))

    def __bool__(self):
        return len(self._fr) > 0

    def __contains__(self, f):
        return f in self._span

    def __iter__(self):
        return iter(self._span)

    def keys(self):
        return self._span.keys()

    def __missing__(self, f):
        a, b = self._span[f]
        v = [(int(i), int(j)) for i, j in self._ev[a:b, 1:3]]
        self[f] = v
        return v

    def __getitem__(self, f):
        if not dict.__contains__(self, f):
            return self.__missing__(f)
        return dict.__getitem__(self, f)

    def items(self):
        return ((f, self[f]) for f in self._span)

    def values(self):
        return (self[f] for f in self._span)


def _group_events(ev: np.ndarray):
    """(frame, i, j) rows -> {frame: [(i, j), ...]} (rows sorted by frame first)."""
    if len(ev) and np.any(np.diff(ev[:, 0]) < 0):
        ev = ev[np.lexsort((ev[:, 2], ev[:, 1], ev[:, 0]))]
    return _LazyPairs(ev)


def build_topology(wrapped: torch.Tensor, symbols, cell, dmax: float, bond_extra: float = 0.45,
                   check_frames=None, dynamic: bool = False, capacity: int = 1 << 16):
    """Topology from frame 0, verified on every frame by the device scan.

    dynamic=False: a bond list that changes along the trajectory raises (the dipolar path indexes pairs from a
    static topology).  dynamic=True: returns (FrameTopologies, two0) instead — the deviating frames are listed by
    the device scan and re-indexed on the host, so that molecules which are not intact in every frame can be
    dropped the way the reference does (SRparse.py:171-176)."""
    A, _, F = wrapped.shape
    two0 = frame_pairs(wrapped, cell, 0, dmax)
    b0 = bonded(two0, symbols, bond_extra)
    topo = Topology(A, two0["i"][b0], two0["j"][b0], symbols)
    if check_frames is None:  # every frame, on the device
        expect = torch.zeros((A, A), dtype=torch.uint8)
        expect[torch.from_numpy(two0["i"][b0]), torch.from_numpy(two0["j"][b0])] = 1
        rc = torch.tensor([SYM2COVRAD[s] for s in symbols], dtype=torch.float64, device=wrapped.device)
        if dynamic:
            # several ranks (every rank holds the whole trajectory): each scans its block of frames, the lists are merged
            from . import dist
            rank, world = dist.rank_world()
            if world > 1:
                lo, hi = dist.shard_range(F, rank, world)
                mine = ops.topology_events(wrapped, rc, bond_extra, expect.to(wrapped.device), cell, dmax, capacity,
                                           frames=(lo, hi))
                parts = dist.all_gather_object(mine)
                ev = np.concatenate([p for p in parts if len(p)] or [mine])
            else:
                ev = ops.topology_events(wrapped, rc, bond_extra, expect.to(wrapped.device), cell, dmax, capacity)
            return FrameTopologies(topo, _group_events(ev), F), two0
        bad, first = ops.topology_check(wrapped, rc, bond_extra, expect.to(wrapped.device), cell, dmax)
        if bad:
            raise RuntimeError("bond topology changes along the trajectory (%d pair-frames differ from frame index 0, "
                               "first at frame index %d): dynpy_b200 indexes molecules from a static topology, the "
                               "reference would re-detect molecules in every frame" % (bad, first))
        return topo, two0
    for f in check_frames:
        two = frame_pairs(wrapped, cell, f, dmax)
        b = bonded(two, symbols, bond_extra)
        if set(zip(two["i"][b].tolist(), two["j"][b].tolist())) != topo.bonds:
            raise RuntimeError("bond topology of frame index %d differs from frame 0: dynpy_b200 indexes molecules "
                               "from a static topology (the reference would re-detect molecules per frame)" % f)
    return topo, two0
