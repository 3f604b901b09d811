"""Which geometry-kernel mode fails at a frame count?   python tools/dbg/dbg_geom_modes.py F"""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 99998
box = synth.water_box(8, seed=3).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
L = box.celldm
w = ops.wrap(box.trajectory(F, chunk=1 << 15, atom_mask=hidx), L)
i, j = np.triu_indices(16, k=1)
pi = torch.tensor(i[:37], dtype=torch.int32).cuda(); pj = torch.tensor(j[:37], dtype=torch.int32).cuda()
for mode in sys.argv[2:] or ("scalar", "vec", "tma"):
    os.environ["DYNPY_DD_GEOM"] = mode
    try:
        out = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), ws_budget=1 << 30)
        torch.cuda.synchronize()
        print(mode, "ok", float(out.abs().sum()))
    except Exception as e:
        print(mode, "FAILED", e)
        break
