"""DD-3 stage at the cfg2 shape (768 atoms x F frames): wrap, build_topology(dynamic=True) (= frame-0 pair table
+ k_topology_check over every frame), select_pairs.   python tools/dbg/probe_topo.py [nmol] [F]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import DDparse, ops, synth, topology
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
box = synth.water_box(nmol, seed=2).to("cuda")
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
dt_sel, (pi, pj, ch) = t(lambda: DDparse.select_pairs_dynamic(box.symbols, ft, 3, "all", None))
A = raw.shape[0]
print("atoms %d frames %d: wrap %.2f ms, build_topology %.2f ms (%.3e atom-pair-frames/s), select_pairs %.2f ms, "
      "deviating frames %d, pairs %d" % (A, F, dt_wrap * 1e3, dt_topo * 1e3, A * (A - 1) / 2 * F / dt_topo, dt_sel * 1e3,
                                         len(ft.frames), len(pi)))
