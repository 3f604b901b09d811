"""Trajectory ingestion with the semantics of the reference's parseMD.py, straight into the
device layout ``[atom][axis][frame]`` the kernels stream (no DataFrame in between).

Mirrors ``PARSE_MD(PD, traj)`` (parseMD.py:11-65):
  * 'xyz'    : blocks of ``nat`` + comment + ``sym x y z`` (parse_xyz, :230-278)
  * 'Tinker' : .arc blocks of ``nat title`` + box line + ``idx sym x y z ...`` (parse_tinker_md, :177-228)
  * printed frames are 1-indexed, ``start_prod``..``end_prod`` inclusive, every ``sample_freq``-th;
    frame id = printed number x ``md_print_freq``; Angstrom -> bohr by ``/0.529177``;
    Fortran ``D`` exponents accepted (d_to_e, :325-329); symbols normalised by normsym (:317-323).
  * 'QE'     : .pos blocks of one header line ``step time`` + ``nat`` lines ``x y z`` (parse_qe_md, :67-153):
    the directory searched is ``traj`` itself, NOT ``traj_dir + traj`` (:22); frame ids are the first token
    of the kept header lines, not printed number x ``md_print_freq`` (:84-92); coordinates are taken as they
    are (no Angstrom -> bohr division); symbols come from ``ParseDynamics.symbols``.  With ``parse_vel`` the
    .vel file is read the same way from columns 1-3 (:126-127); the reference itself raises there as soon as
    more than one frame is kept (``velatom.loc[:, 'symbol'] = symbols``, :131, a length mismatch), so that
    branch follows the evident intent and is pinned by no reference output.
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


def _read_blocks(path, nat, header_lines, cols, printed, scale_div=1.0, pin=False, header_ids=None):
    """(symbols, xyz[F, A, 3] / scale_div) for the 1-indexed printed frames, parsed by the native
    multi-threaded reader of the C-ABI library (csrc/ingest.cpp: mmap, parallel newline index,
    std::from_chars).  `pin` puts the array in page-locked memory for the H2D copy.  cols[0] = -1: no symbol
    token (symbols come back as None); header_ids: float64 array [F] that receives the first header token of
    every kept block."""
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
    if header_ids is not None:
        assert header_ids.dtype == np.float64 and header_ids.shape == (F,) and header_ids.flags.c_contiguous
        rc = lib.dynpy_parse_md_blocks_hdr_host(os.fsencode(path), nat, header_lines, cols[0], printed.ctypes.data, F,
                                                float(scale_div), host.data_ptr(), C.cast(sym, C.c_void_p),
                                                header_ids.ctypes.data, 0)
    else:
        rc = lib.dynpy_parse_md_blocks_host(os.fsencode(path), nat, header_lines, cols[0], printed.ctypes.data, F,
                                            float(scale_div), host.data_ptr(), C.cast(sym, C.c_void_p), 0)
    _lib.check(rc, "dynpy_parse_md_blocks_host")
    if cols[0] < 0:
        return None, host
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


def _load_frames(path, nat, header_lines, cols, printed, device, scale_div=0.529177, header_ids=None):
    """(symbols, coords [A, 3, F] on `device`) of the printed frames.  With several ranks every rank parses only
    its block of frames (the text is the expensive part), copies it over its own host link and the blocks are
    exchanged GPU to GPU (one all-gather), so that every rank ends up with the whole trajectory.
    header_ids (float64 [F], optional) is filled on every rank."""
    import torch.distributed as tdist

    from . import dist
    rank, world = dist.rank_world()
    F = len(printed)
    if world == 1 or device.type != "cuda" or F < world:
        sym, xyz = _read_blocks(path, nat, header_lines, cols, printed, scale_div, pin=True, header_ids=header_ids)
        # [F][A][3] (pinned host) -> device -> the kernels' [A][3][F] layout (transpose on the device)
        return sym, xyz.to(device, non_blocking=True).permute(1, 2, 0).contiguous()
    cuts = [(F * r) // world for r in range(world + 1)]
    mine_ids = np.empty(cuts[rank + 1] - cuts[rank], dtype=np.float64) if header_ids is not None else None
    sym, mine = _read_blocks(path, nat, header_lines, cols, printed[cuts[rank]:cuts[rank + 1]], scale_div, pin=True,
                             header_ids=mine_ids)
    full = torch.empty((F, nat, 3), dtype=torch.float64, device=device)
    part = full[cuts[rank]:cuts[rank + 1]]
    part.copy_(mine, non_blocking=True)
    dist.gather_blocks(full, cuts, part)
    box = [(sym, mine_ids)]
    if header_ids is not None:
        parts = [None] * world
        tdist.all_gather_object(parts, mine_ids)
        header_ids[:] = np.concatenate(parts)
    tdist.broadcast_object_list(box, src=0)   # symbols of the first printed frame, as in the single-process run
    return box[0][0], full.permute(1, 2, 0).contiguous()


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
    scale_div, header_ids, vcols = 0.529177, None, None
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
        traj_dir = traj                       # parseMD.py:22 passes traj_dir=traj (PD.traj_dir is not prepended)
        try:
            name = [f for f in os.listdir(traj_dir) if ".pos" in f][0]
        except Exception:
            print("Did not find .pos trajectory file at " + traj_dir + " for parsing QE dynamics. Is this what you wanted?")
            sys.exit(2)
        if len(PD.symbols) != PD.nat:
            # the reference sizes the blocks by len(symbols) (parseMD.py:79); nat is a required key of the CLI
            raise ValueError("ParseDynamics.symbols lists %d atoms, nat = %d" % (len(PD.symbols), PD.nat))
        header, cols, vcols = 1, (-1, 0, 1, 2), (0, 1, 2, 3)
        scale_div = 1.0                       # QE coordinates are used as given (bohr)
        header_ids = np.empty(len(printed), dtype=np.float64)
    else:
        print("MD_format not provided or not recognized. Implemented formats are 'QE', 'Tinker', and 'xyz'. "
              "Do you need to parse MD trajectories?")
        sys.exit(2)
    sym, pos = _load_frames(os.path.join(traj_dir, name), PD.nat, header, cols, printed, device, scale_div, header_ids)
    if fmt == "QE":
        sym = [normsym(s) for s in PD.symbols]
        frames = header_ids.astype(np.int64)  # the step ids of the kept header lines (parseMD.py:84-92)
        if not np.array_equal(frames.astype(np.float64), header_ids):
            raise ValueError("non-integer step ids in the .pos headers of %s" % name)
    else:
        frames = printed.astype(np.int64) * int(PD.md_print_freq)
    vel = None
    if PD.parse_vel:
        vname = _find(traj_dir, ".vel", ".vel")
        _, vel = _load_frames(os.path.join(traj_dir, vname), PD.nat, header, vcols or cols, printed, device, scale_div)
    return Trajectory(pos=pos, symbols=sym, frames=frames, vel=vel)
