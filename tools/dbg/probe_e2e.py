"""Stage timers of bench.py's end-to-end leg at cfg2 (one GPU): where the time between `value` and `e2e` goes."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import DDparse, ddrax, ops, synth, topology
dev = torch.device("cuda")
nmol, F = 256, 100000
box = synth.water_box(nmol, seed=2).to(dev)
L = box.celldm; cell = (L, L, L)
full = box.trajectory(F, chunk=2048)
host = torch.empty(full.shape, dtype=torch.float64, pin_memory=True); host.copy_(full); del full
staging = torch.empty(host.shape, dtype=torch.float64, device=dev)
wrapped = torch.empty_like(staging)
tdev = torch.arange(1, F + 1, device=dev, dtype=torch.float64) * 0.001
M = ops.fft_length(F)
host_G = torch.empty((7, F), dtype=torch.float64, pin_memory=True)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return time.perf_counter()
def step():
    t = time.perf_counter()
    staging.copy_(host, non_blocking=True); t = tick("h2d", t)
    ops.wrap(staging, L, out=wrapped); t = tick("wrap", t)
    ft, _ = topology.build_topology(wrapped, box.symbols, cell, L, bond_extra=0.45, dynamic=True, capacity=1 << 22); t = tick("build_topology", t)
    sp_i, sp_j = DDparse.select_pairs(box.symbols, ft.base, 3, "all", None); t = tick("select_pairs", t)
    qi, qj = torch.from_numpy(sp_i).to(dev), torch.from_numpy(sp_j).to(dev); t = tick("pairs h2d", t)
    G = torch.zeros((7, F), dtype=torch.float64, device=dev)
    spec = ops.dd_pairs_to_spectrum(wrapped, qi, qj, cell, M=M); t = tick("transforms", t)
    G += ops.ifft_acf(spec, F, 1.0 / (float(M) * float(F))); G *= 2.0 / (2 * nmol); t = tick("ifft", t)
    r = ddrax.dipolar(ddrax.spec_dens_extr_narr(tdev, G, "cartwright"), "1H", "1H"); t = tick("simpson+rate", t)
    host_G.copy_(G, non_blocking=True); t = tick("d2h", t)
    return len(ft.changes)
step(); T.clear()
n = 3
t0 = time.perf_counter()
for _ in range(n): dv = step()
tot = (time.perf_counter() - t0) / n
print("deviating frames", dv, " total %.1f ms" % (tot * 1e3))
for k, v in T.items(): print("  %-16s %8.2f ms" % (k, v / n * 1e3))
