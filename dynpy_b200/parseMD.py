"""Trajectory ingestion with the semantics of the reference's parseMD.py, straight into the
device layout ``[atom][axis][frame]`` the kernels stream (no DataFrame in between).

Mirrors ``PARSE_MD(PD, traj)`` (parseMD.py:11-65):
  * 'xyz'    : blocks of ``nat`` + comment + ``sym x y z`` (parse_xyz, :230-278)
  * 'Tinker' : .arc blocks of ``nat title`` + box line + ``idx sym x y z ...`` (parse_tinker_md, :177-228)
  * printed frames are 1-indexed, ``start_prod``..``end_prod`` inclusive, every ``sample_freq``-th;
    frame id = printed number x ``md_print_freq``; Angstrom -> bohr by ``/0.529177``;
    Fortran ``D`` exponents accepted (d_to_e, :325-329); symbols normalised by normsym (:317-323).
  * 'QE'     : the reference's reader returns nothing (parse_qe_md has no return, :67-134), so the
    reference itself cannot run this format; we raise a clear error instead of guessing.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Trajectory:
    pos: torch.Tensor          # [A, 3, F] float64, bohr, raw (unwrapped-as-given), on `device`
    symbols: list              # [A]
    frames: np.ndarray         # [F] int64 frame ids (printed number x md_print_freq)
    vel: torch.Tensor | None = None

    @property
    def nat(self):
        return self.pos.shape[0]

    @property
    def nframes(self):
        return self.pos.shape[2]


def normsym(sym):
    sym = str(sym)
    if len(sym) > 1:
        if sym[1].isupper() or sym[1].isnumeric():
            sym = sym[0]
    return sym


def _select_frames(start_prod, end_prod, sample_freq):
    ref = np.arange(start_prod, end_prod + sample_freq, sample_freq)
    sel = ref[ref <= end_prod]
    if len(sel) != len(ref):
        print("Warning: (end_prod - start_prod) is not a multiple of sample_freq; the reference fails on "
              "such input (parseMD.py:240-244). Using printed frames %d..%d." % (sel[0], sel[-1]))
    return sel


def _read_blocks(path, nat, header_lines, cols, printed, scale_div=1.0, pin=False):
    """(symbols, xyz[F, A, 3] / scale_div) for the 1-indexed printed frames, parsed by the native
    multi-threaded reader of the C-ABI library (csrc/ingest.cpp: mmap, parallel newline index,
    std::from_chars).  `pin` puts the array in page-locked memory for the H2D copy."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    printed = np.ascontiguousarray(np.asarray(printed, dtype=np.int64))
    F = len(printed)
    if pin and torch.cuda.is_available():
        host = torch.empty((F, nat, 3), dtype=torch.float64, pin_memory=True)
    else:
        host = torch.empty((F, nat, 3), dtype=torch.float64)
    sym = C.create_string_buffer(8 * nat)
    rc = lib.dynpy_parse_md_blocks_host(os.fsencode(path), nat, header_lines, cols[0], printed.ctypes.data, F,
                                        float(scale_div), host.data_ptr(), C.cast(sym, C.c_void_p), 0)
    _lib.check(rc, "dynpy_parse_md_blocks_host")
    raw = sym.raw
    symbols = [normsym(raw[8 * a:8 * a + 8].split(b"\0", 1)[0].decode()) for a in range(nat)]
    return symbols, host


def _read_blocks_py(path, nat, header_lines, cols, printed):
    """Pure-Python restatement of the same block reader (test checker for the native one):
    (symbols, xyz[F, A, 3] in file units) for the 1-indexed printed frames."""
    blk = nat + header_lines
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    need = int(printed[-1]) * blk
    if len(lines) < need:
        raise ValueError("%s holds fewer than %d printed frames of %d atoms" % (path, printed[-1], nat))
    out = np.empty((len(printed), nat, 3), dtype=np.float64)
    symbols = None
    for k, p in enumerate(printed):
        s = (int(p) - 1) * blk + header_lines
        toks = [ln.split() for ln in lines[s:s + nat]]
        if k == 0:
            symbols = [normsym(t[cols[0]]) for t in toks]
        flat = [t[c] for t in toks for c in cols[1:]]
        try:
            arr = np.array(flat, dtype=np.float64)
        except ValueError:
            arr = np.array([float(v.replace("D", "E").replace("d", "e")) for v in flat], dtype=np.float64)
        out[k] = arr.reshape(nat, 3)
    return symbols, out


def _load_frames(path, nat, header_lines, cols, printed, device):
    """(symbols, coords [A, 3, F] on `device`) of the printed frames.  With several ranks every rank parses only
    its block of frames (the text is the expensive part), copies it over its own host link and the blocks are
    exchanged GPU to GPU, so that every rank ends up with the whole trajectory."""
    import torch.distributed as tdist

    from . import dist
    rank, world = dist.rank_world()
    F = len(printed)
    if world == 1 or device.type != "cuda" or F < world:
        sym, xyz = _read_blocks(path, nat, header_lines, cols, printed, 0.529177, pin=True)
        # [F][A][3] (pinned host) -> device -> the kernels' [A][3][F] layout (transpose on the device)
        return sym, xyz.to(device, non_blocking=True).permute(1, 2, 0).contiguous()
    cuts = [(F * r) // world for r in range(world + 1)]
    sym, mine = _read_blocks(path, nat, header_lines, cols, printed[cuts[rank]:cuts[rank + 1]], 0.529177, pin=True)
    full = torch.empty((F, nat, 3), dtype=torch.float64, device=device)
    full[cuts[rank]:cuts[rank + 1]].copy_(mine, non_blocking=True)
    for r in range(world):
        tdist.broadcast(full[cuts[r]:cuts[r + 1]], src=r)
    box = [sym]
    tdist.broadcast_object_list(box, src=0)   # symbols of the first printed frame, as in the single-process run
    return box[0], full.permute(1, 2, 0).contiguous()


def _find(traj_dir, token, what):
    try:
        return [f for f in os.listdir(traj_dir) if token in f][0]
    except Exception:
        print("Did not find %s trajectory file at %s for parsing dynamics. Is this what you wanted?" % (what, traj_dir))
        sys.exit(2)


def PARSE_MD(PD, traj, device=None):
    """-> Trajectory on `device` (default: current CUDA device)."""
    if 'parse_vel' not in PD.__dict__.keys():
        PD.parse_vel = False
    if 'sample_freq' not in PD.__dict__.keys():
        PD.sample_freq = 1
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    fmt = PD.MD_format
    traj_dir = PD.traj_dir + traj
    printed = _select_frames(PD.start_prod, PD.end_prod, PD.sample_freq)
    if fmt == "xyz":
        name = _find(traj_dir, ".xyz", ".xyz")
        header, cols = 2, (0, 1, 2, 3)
    elif fmt == "Tinker":
        name = _find(traj_dir, "arc", ".arc")
        print("reading " + name)
        header, cols = 2, (1, 2, 3, 4)
    elif fmt == "QE":
        if 'symbols' not in PD.__dict__.keys():
            print("Error: Missing required input variable symbols in class ParseDynamics for parsing QE.")
            sys.exit(2)
        raise NotImplementedError("MD_format 'QE': the reference's parse_qe_md returns nothing (parseMD.py:67-134), "
                                  "so there is no behaviour to reproduce; convert the .pos file to xyz")
    else:
        print("MD_format not provided or not recognized. Implemented formats are 'QE', 'Tinker', and 'xyz'. "
              "Do you need to parse MD trajectories?")
        sys.exit(2)
    sym, pos = _load_frames(os.path.join(traj_dir, name), PD.nat, header, cols, printed, device)
    frames = printed.astype(np.int64) * int(PD.md_print_freq)
    vel = None
    if PD.parse_vel:
        vname = _find(traj_dir, ".vel", ".vel")
        _, vel = _load_frames(os.path.join(traj_dir, vname), PD.nat, header, cols, printed, device)
    return Trajectory(pos=pos, symbols=sym, frames=frames, vel=vel)
