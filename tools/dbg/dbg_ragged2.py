import sys, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, topology, DDparse, ddrax
from oracle import oracle_np as orc
box = synth.water_box(27, seed=12, drift=0.01)
F = 300
traj = box.frames(0, F)
L = box.celldm
cell = (L,)*3
wrapped = ops.wrap(traj.cuda().contiguous(), L)
w = wrapped.cpu().numpy()
rmax = 9.5
topo, _ = topology.build_topology(wrapped, box.symbols, cell, rmax)
pi, pj = DDparse.select_pairs(box.symbols, topo, 3, "inter", None)
cnt = ops.dd_pair_count(wrapped, torch.from_numpy(pi).cuda(), torch.from_numpy(pj).cuda(), cell, rmax).cpu().numpy()
# numpy presence
def pair_ref(i, j):
    d = np.empty((3, F))
    for ax in range(3):
        c = np.stack([(w[i, ax] + aa * L) - w[j, ax] for aa in (-1, 0, 1)])
        k = np.argmin(c ** 2, axis=0)
        d[ax] = c[k, np.arange(F)]
    r2 = d[0]**2 + d[1]**2 + d[2]**2
    pres = r2 < rmax**2
    ref = np.zeros((7, F))
    if pres.any():
        dr = np.sqrt(r2[pres])
        y20, y21, y22, r3, f0, f1, f2 = orc.dd_words(d[0][pres], d[1][pres], d[2][pres], dr)
        fr = np.nonzero(pres)[0]
        for c_, v in enumerate([y20, y21, y22, r3 - r3.mean(), f0, f1, f2]):
            ref[c_, fr] = orc.wiener_khinchin(v)
    return pres.sum(), ref
refs = [pair_ref(int(a), int(b)) for a, b in zip(pi, pj)]
rc = np.array([r[0] for r in refs])
print("count mismatches:", (rc != cnt).sum())
full = cnt == F; part = (cnt > 0) & ~full
Gfull_ref = sum(refs[k][1] for k in np.nonzero(full)[0])
Gpart_ref = sum(refs[k][1] for k in np.nonzero(part)[0])
ti = lambda m: torch.from_numpy(pi[m]).cuda().contiguous(); tj = lambda m: torch.from_numpy(pj[m]).cuda().contiguous()
M = ops.fft_length(F)
spec = ops.dd_pairs_to_spectrum(wrapped, ti(full), tj(full), cell, M=M)
Gfull = ops.ifft_acf(spec, F, 1.0/(M*F)).cpu().numpy()
Gpart = ops.dd_pairs_ragged_to_acf(wrapped, ti(part), tj(part), cell, rmax).cpu().numpy()
ne = lambda a, b: np.max(np.abs(a-b))/np.max(np.abs(b))
print("full err", [ "%.1e" % ne(Gfull[c], Gfull_ref[c]) for c in range(7)])
print("part err", [ "%.1e" % ne(Gpart[c], Gpart_ref[c]) for c in range(7)])
# batches of different sizes
idx = np.nonzero(part)[0]
for nb in (1, 2, 3, 8, 50, 200, 600, len(idx)):
    sel = idx[:nb]
    G = ops.dd_pairs_ragged_to_acf(wrapped, torch.from_numpy(pi[sel]).cuda(), torch.from_numpy(pj[sel]).cuda(), cell, rmax).cpu().numpy()
    R = sum(refs[k][1] for k in sel)
    print(nb, "%.1e" % max(ne(G[c], R[c]) for c in range(7)))
