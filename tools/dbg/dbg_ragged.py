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
part = np.nonzero((cnt > 0) & (cnt < F))[0]
print("pairs", len(pi), "partial", len(part), "full", (cnt == F).sum())
bad = 0
for idx in part[:400]:
    i, j = int(pi[idx]), int(pj[idx])
    G = ops.dd_pairs_ragged_to_acf(wrapped, torch.tensor([i], dtype=torch.int32).cuda(), torch.tensor([j], dtype=torch.int32).cuda(), cell, rmax).cpu().numpy()
    # numpy per pair
    d = np.empty((3, F)); 
    for ax in range(3):
        c = np.stack([(w[i, ax] + aa * L) - w[j, ax] for aa in (-1, 0, 1)])
        k = np.argmin(c ** 2, axis=0)
        d[ax] = c[k, np.arange(F)]
    r2 = d[0]**2 + d[1]**2 + d[2]**2
    pres = r2 < rmax**2
    dr = np.sqrt(r2[pres])
    y20, y21, y22, r3, f0, f1, f2 = orc.dd_words(d[0][pres], d[1][pres], d[2][pres], dr)
    ref = np.zeros((7, F))
    fr = np.nonzero(pres)[0]
    for c_, v in enumerate([y20, y21, y22, r3 - r3.mean(), f0, f1, f2]):
        ref[c_, fr] = orc.wiener_khinchin(v)
    err = [np.max(np.abs(G[c_] - ref[c_])) / max(np.max(np.abs(ref[c_])), 1e-300) for c_ in range(7)]
    if max(err) > 1e-9:
        bad += 1
        if bad < 6:
            print("pair", i, j, "Np", pres.sum(), "cnt", cnt[idx], "err", ["%.1e" % e for e in err], "first present", fr[:5], "gaps at", np.nonzero(~pres)[0][:8])
print("bad", bad, "of", min(400, len(part)))
