"""Host-side stage timers of topology.build_topology(dynamic=True) + select_pairs at the cfg2 shape (768 atoms x 1e5 frames)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import DDparse, ops, synth, topology
from dynpy_b200.elements import SYM2COVRAD
dev = torch.device("cuda")
box = synth.water_box(256, seed=2).to(dev)
F = 100000
L = box.celldm; cell = (L, L, L)
w = ops.wrap(box.trajectory(F, chunk=2048), L)
sym = box.symbols
def T(name, fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); print("  %-34s %8.2f ms" % (name, (time.perf_counter() - t0) / n * 1e3)); return r
two0 = T("frame_pairs(frame 0, dmax = L)", lambda: topology.frame_pairs(w, cell, 0, L))
b0 = T("bonded", lambda: topology.bonded(two0, sym, 0.45))
topo = T("Topology (union-find)", lambda: topology.Topology(w.shape[0], two0["i"][b0], two0["j"][b0], sym))
def mk():
    e = torch.zeros((w.shape[0], w.shape[0]), dtype=torch.uint8)
    e[torch.from_numpy(two0["i"][b0]), torch.from_numpy(two0["j"][b0])] = 1
    rc = torch.tensor([SYM2COVRAD[s] for s in sym], dtype=torch.float64, device=dev)
    return e.to(dev), rc
expect, rc = T("expect matrix + radii to device", mk)
ev = T("topology_events (scan + sort + D2H)", lambda: ops.topology_events(w, rc, 0.45, expect, cell, L, 1 << 22), 3)
print("   events:", len(ev))
ft = T("FrameTopologies + grouping", lambda: topology.FrameTopologies(topo, topology._group_events(ev), F))
T("len(ft.changes)", lambda: len(ft.changes))
T("select_pairs", lambda: DDparse.select_pairs(sym, ft.base, 3, "all", None))
T("build_topology total", lambda: topology.build_topology(w, sym, cell, L, bond_extra=0.45, dynamic=True, capacity=1 << 22), 3)
