import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth
box = synth.water_box(8, seed=3).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
L = box.celldm
i, j = np.triu_indices(16, k=1)
for F in [int(a) for a in sys.argv[1:]]:
    w = ops.wrap(box.trajectory(F, chunk=1 << 15, atom_mask=hidx), L)
    for P in (2, 37):
        pi = torch.tensor(i[:P], dtype=torch.int32).cuda(); pj = torch.tensor(j[:P], dtype=torch.int32).cuda()
        for ws in (1 << 30, None):
            try:
                out = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), ws_budget=ws)
                torch.cuda.synchronize()
                print(F, P, ws, "ok", float(out.abs().sum()))
            except Exception as e:
                print(F, P, ws, "FAILED", str(e)[-60:])
