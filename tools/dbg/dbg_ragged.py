"""Throughput of the ragged dipolar path (pairs that leave rmax) next to the gapless path, cfg2 frame count."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, ddrax
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
F = 100000
box = synth.water_box(256, seed=2, drift=0.002).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")[:128]
w = ops.wrap(box.trajectory(F, chunk=4096, atom_mask=hidx), box.celldm)
L = box.celldm
i, j = np.triu_indices(128, k=1)
rng = np.random.default_rng(0)
sel = rng.permutation(len(i))[:P]
pi = torch.tensor(i[sel], dtype=torch.int32).cuda(); pj = torch.tensor(j[sel], dtype=torch.int32).cuda()
cell = (L, L, L)
for rmax in (15.0, 25.0, L):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        G, hit, ns, info = ddrax.pair_acf_sum(w, pi, pj, cell, rmax, shard=False)
        torch.cuda.synchronize(); dt = time.time() - t0
    print("rmax %.1f (L=%.1f): %s  %.3f s  %.3e pair-frames/s (all frames counted)" % (rmax, L, {k: info[k] for k in ("full", "partial", "absent")}, dt, P * F / dt))
