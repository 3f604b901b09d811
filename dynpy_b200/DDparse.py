"""Dipolar driver — mirror of the reference's DDparse.py (DD_module_main :17-132,
prep_DD_uni1 :133-169) on top of the CUDA path.  Same parameter classes, prompts, stdout
result lines and CSV files (``<traj_dir><traj>/avg-acf.csv``, ``<traj_dir>DD-results.csv``)."""
from __future__ import annotations

import sys
import time

import numpy as np
import pandas as pd
import torch

from . import ddrax, dist, ops, parseMD, topology
from .config import default_simpson_mode, dump_stages, stage
from .helper import which_trajs


def _eligible(symbols, nat_per_mol, analyte_label):
    """(i, j, sel): all label pairs i < j and the symbol / analyte filter of DDparse.py:53-56."""
    A = len(symbols)
    sym = np.array(symbols)
    i, j = np.triu_indices(A, k=1)
    if analyte_label:  # label 0 is falsy in the reference as well (DDparse.py:53)
        mai = np.arange(A) % nat_per_mol
        sel = (mai[i] == analyte_label) & (sym[j] == 'H')
    else:
        sel = (sym[i] == 'H') & (sym[j] == 'H')
    return i, j, sel


def select_pairs(symbols, topo, nat_per_mol, pair_type, analyte_label):
    """Pair rows the reference keeps (DDparse.py:53-60): H–H (or analyte_label–H) with
    label0 < label1, intra / inter / all by molecule membership.  Returns int32 arrays."""
    i, j, sel = _eligible(symbols, nat_per_mol, analyte_label)
    m = topo.mol_rank
    if pair_type == 'intra':
        sel &= m[i] == m[j]
    elif pair_type == 'inter':
        sel &= m[i] != m[j]
    return i[sel].astype(np.int32), j[sel].astype(np.int32)


def select_pairs_dynamic(symbols, ft, nat_per_mol, pair_type, analyte_label):
    """The same for a trajectory whose bond list is not the same in every frame (ft: FrameTopologies).  The
    reference filters pair ROWS by the molecule ids of their own frame, so a pair can be intramolecular in some
    frames and intermolecular in others.  Returns (pi, pj) for the pairs whose status never changes and
    (mi, mj, mask [K, F] uint8) for those that do, mask = frames in which the row is kept."""
    i, j, sel = _eligible(symbols, nat_per_mol, analyte_label)
    base_same = ft.base.mol_rank[i] == ft.base.mol_rank[j]
    if pair_type not in ('intra', 'inter') or not ft.changes:
        pi, pj = select_pairs(symbols, ft.base, nat_per_mol, pair_type, analyte_label)
        return pi, pj, None
    want_same = pair_type == 'intra'
    changed = np.zeros(len(i), dtype=bool)
    same_f = {}
    for f, t in ft.frames.items():
        same_f[f] = t.mol_rank[i] == t.mol_rank[j]
        changed |= same_f[f] != base_same
    changed &= sel
    static = sel & (base_same == want_same) & ~changed
    mask = np.repeat((base_same[changed] == want_same)[:, None], ft.nframes, axis=1)
    for f, sf in same_f.items():
        mask[:, f] = sf[changed] == want_same
    return (i[static].astype(np.int32), j[static].astype(np.int32),
            (i[changed].astype(np.int32), j[changed].astype(np.int32), mask.astype(np.uint8)))


def dd_trajectory(tr, PD, DD, pair_type, analyte_label, analyte_species, simpson_mode):
    """One trajectory -> (acf_avg DataFrame, rate, tau_c, F0, info)."""
    try:
        dmax = DD.rmax
    except AttributeError:
        dmax = 15
    cell = (float(PD.celldm),) * 3
    with stage("wrap + bonds/molecules of every frame (DD-1, DD-3)"):
        wrapped = ops.wrap(tr.pos, cell[0])
        ft, _ = topology.build_topology(wrapped, tr.symbols, cell, dmax, bond_extra=0.45, dynamic=True)
        ft.base.classify_exact(DD.identifier)  # raises like Molecule.classify when the formula is absent
    with stage("pair selection (DD-3)"):
        nat_per_mol = topology.atoms_per_formula(DD.identifier)
        pi, pj, changing = select_pairs_dynamic(tr.symbols, ft, nat_per_mol, pair_type, analyte_label)
        dev = wrapped.device
        if changing is not None:
            changing = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in changing)
        pi_d, pj_d = torch.from_numpy(pi).to(dev), torch.from_numpy(pj).to(dev)
    with stage("pair vectors, Y2m words, transforms, sum over pairs (DD-2, DD-4..6)"):
        G, hit, n_spins, info = ddrax.pair_acf_sum(wrapped, pi_d, pj_d, cell, dmax, masked=changing)
    if n_spins == 0:
        raise ValueError("no pair of the requested kind is ever inside rmax=%r" % (dmax,))
    with stage("Simpson + rates (DD-7)"):
        G = (G[:, hit] * (2.0 / n_spins)).contiguous()
        tdev = (torch.from_numpy(tr.frames).to(dev)[hit].to(torch.float64) * float(PD.timestep)).contiguous()
        EN = ddrax.spec_dens_extr_narr(tdev, G, simpson_mode)
        rax, tc, F = ddrax.dipolar(EN, symbol1=analyte_species, symbol2='1H')
    with stage("G(tau) to host table"):
        acf_avg = pd.DataFrame({'time': tdev.cpu().numpy()})
        Gh = G.cpu().numpy()
        for c, name in enumerate(ddrax.G_COLUMNS):
            acf_avg[name] = Gh[c]
    return acf_avg, rax, tc, F, info


def DD_module_main(PD, DD):
    outer_start_time = time.time()
    analyte_label = getattr(DD, 'analyte_label', None)
    analyte_species = getattr(DD, 'analyte_species', '1H')
    try:
        pair_type = DD.pair_type
        if pair_type.casefold() not in ['intra', 'inter', 'all']:
            pair_type = 'all'
            print("Inter- and intramolecular pairs will be considered.\n")
    except AttributeError:
        pair_type = 'all'
    simpson_mode = getattr(DD, 'simpson_mode', None) or default_simpson_mode()
    rank, world = dist.rank_world()
    # rank 0 asks; its decision (and the trajectory list it may have filled in) goes to every rank, so that
    # an "n" stops the whole job instead of leaving the other ranks in a barrier
    decision = None
    if rank == 0:
        try:
            which_trajs(PD)
            decision = ("go", list(PD.trajs))
        except SystemExit as e:
            decision = ("exit", e.code)
        except EOFError:
            decision = ("eof", None)
    decision = dist.broadcast_object(decision)
    if decision[0] == "exit":
        dist.shutdown()
        sys.exit(decision[1])
    if decision[0] == "eof":
        dist.shutdown()
        raise EOFError("EOF when reading a line")
    PD.trajs = decision[1]
    say = print if rank == 0 else (lambda *a, **k: None)   # user-facing lines once, not once per rank
    res = {}
    for t, traj in enumerate(PD.trajs):
        inner_start_time = time.time()
        with stage("trajectory text -> device [A][3][F] (parse + H2D + transpose)"):
            tr = parseMD.PARSE_MD(PD, traj)
        say("computing two-atom distances...")
        say("computing dipolar quantities...")
        say("computing TCFs...")
        acf_avg, rax, tc, F, info = dd_trajectory(tr, PD, DD, pair_type, analyte_label, analyte_species, simpson_mode)
        torch.cuda.synchronize()
        say("--- {t:.2f} seconds ---".format(t=time.time() - inner_start_time))
        if rank == 0:
            with stage("avg-acf.csv"):
                acf_avg.to_csv(PD.traj_dir + traj + '/avg-acf.csv')
        say("1/T1/nH = {r:.3e} Hz, tau_c = {t:.3e} ps, <F(0)^2> = {f:.3e} au^-6".format(r=rax, t=tc, f=F))
        say("For intramolecular contrib. in heteronuclear systems (e.g. C--H relaxation of 13C), multiply 1/T1/nH "
              "by number of bonded spins to obtain 1/T1. E.g. for C-13 in methyl group, nH=3")
        res[traj] = [rax, tc, F]
    res_df = pd.DataFrame(res).T
    res_df.columns = ["1/T1/nH", "tau_c", "<F(0)^2>"]
    if rank == 0:
        res_df.to_csv(PD.traj_dir + 'DD-results.csv')
        dump_stages({"total_s": time.time() - outer_start_time, "pairs": info["pairs"], "frames": int(tr.nframes),
                     "atoms": int(tr.nat), "full": info["full"], "partial": info["partial"]})
    return res_df
