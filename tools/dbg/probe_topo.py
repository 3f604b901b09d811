"""DD-3 stage at the cfg2 shape (768 atoms x F frames): wrap, build_topology(dynamic=True) (= frame-0 pair table
+ k_topology_check over every frame), select_pairs.   python tools/dbg/probe_topo.py [nmol] [F]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import DDparse, ops, synth, topology
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
import os
box = synth.water_box(nmol, seed=2, trans_amp=float(os.environ.get("TRANS_AMP", "0.03"))).to("cuda")
raw = box.trajectory(F, chunk=4096)
L = box.celldm
cell = (L, L, L)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / reps, out
dt_wrap, w = t(lambda: ops.wrap(raw, L))
dt_topo, (ft, two0) = t(lambda: topology.build_topology(w, box.symbols, cell, L, bond_extra=0.45, dynamic=True))
# breakdown: the device scan alone, then the host re-indexing of the deviating frames
from dynpy_b200.elements import SYM2COVRAD as _RC
_two0 = topology.frame_pairs(w, cell, 0, L)
_b0 = topology.bonded(_two0, box.symbols, 0.45)
_exp = torch.zeros((raw.shape[0], raw.shape[0]), dtype=torch.uint8)
_exp[torch.from_numpy(_two0["i"][_b0]), torch.from_numpy(_two0["j"][_b0])] = 1
_exp = _exp.cuda()
_rc = torch.tensor([_RC[s] for s in box.symbols], dtype=torch.float64, device="cuda")
dt_scan, ev = t(lambda: ops.topology_events(w, _rc, 0.45, _exp, cell, L))
dt_fp, _ = t(lambda: topology.frame_pairs(w, cell, 0, L))
print("  device scan (topology_events) %.2f ms = %.3e atom-pair-frames/s, %d events; frame-0 pair table %.2f ms"
      % (dt_scan * 1e3, raw.shape[0] * (raw.shape[0] - 1) / 2 * F / dt_scan, len(ev), dt_fp * 1e3))
dt_sel, (pi, pj, ch) = t(lambda: DDparse.select_pairs_dynamic(box.symbols, ft, 3, "all", None))
A = raw.shape[0]
print("atoms %d frames %d: wrap %.2f ms, build_topology %.2f ms (%.3e atom-pair-frames/s), select_pairs %.2f ms, "
      "deviating frames %d, pairs %d" % (A, F, dt_wrap * 1e3, dt_topo * 1e3, A * (A - 1) / 2 * F / dt_topo, dt_sel * 1e3,
                                         len(ft.changes), len(pi)))

# cross-check of the device scan against a brute-force torch evaluation on the first frames
from dynpy_b200.elements import SYM2COVRAD
nchk = 256
sub = w[:, :, :nchk].contiguous()
two0 = topology.frame_pairs(sub, cell, 0, L)
b0 = topology.bonded(two0, box.symbols, 0.45)
expect = torch.zeros((A, A), dtype=torch.uint8)
expect[torch.from_numpy(two0["i"][b0]), torch.from_numpy(two0["j"][b0])] = 1
rc = torch.tensor([SYM2COVRAD[s] for s in box.symbols], dtype=torch.float64, device="cuda")
bad, first = ops.topology_check(sub, rc, 0.45, expect.cuda(), cell, L)
thr = rc[:, None] + rc[None, :] + 0.45
iu = torch.triu(torch.ones((A, A), dtype=torch.bool, device="cuda"), 1)
ref_bad = 0
for f0 in range(0, nchk, 32):
    x = sub[:, :, f0:f0 + 32]                                  # [A,3,T]
    d = x[:, None] - x[None, :]                                 # [A,A,3,T]
    d = d - L * torch.round(d / L)
    r = torch.sqrt((d * d).sum(2))                              # [A,A,T]
    bond = r <= thr[:, :, None]
    ref_bad += int(((bond != expect.cuda()[:, :, None].bool()) & iu[:, :, None]).sum())
print("topology_check on %d frames: device %d mismatching pair-frames, torch brute force %d" % (nchk, bad, ref_bad))
