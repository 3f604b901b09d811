"""Fused dipolar call on P pairs x F frames (64 H atoms): per-class device time of this process's kernel
selection (env knobs are read once per process).   python tools/dbg/probe_dd.py P F [reps] [ws_gb]"""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ws = float(sys.argv[4]) if len(sys.argv) > 4 else 4.0
box = synth.water_box(32, seed=2).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
w = ops.wrap(box.trajectory(F, chunk=1 << 16, atom_mask=hidx), box.celldm)
L = box.celldm
i, j = np.triu_indices(64, k=1)
rep = (P + len(i) - 1) // len(i)
i, j = np.tile(i, rep)[:P], np.tile(j, rep)[:P]
pi = torch.tensor(i, dtype=torch.int32).cuda(); pj = torch.tensor(j, dtype=torch.int32).cuda()
lib = _lib.load()
M = ops.fft_length(F)
spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M, ws_budget=int(ws * (1 << 30)))
best = None
for r in range(reps):
    lib.dynpy_profile_enable(1); lib.dynpy_profile_read(None, None)
    torch.cuda.synchronize(); t0 = time.time()
    spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M, ws_budget=int(ws * (1 << 30)))
    torch.cuda.synchronize(); dt = time.time() - t0
    ms = (ctypes.c_double * 4)(); ln = (ctypes.c_int64 * 4)()
    lib.dynpy_profile_read(ms, ln); lib.dynpy_profile_enable(0)
    if best is None or dt < best[0]:
        best = (dt, list(ms[:3]))
G = ops.ifft_acf(spec, F, 1.0 / (M * F))
print("P %d F %d M %d  %.4f s  %.3e pf/s  geom/cols/rows ms %s  G0 %s" % (
    P, F, M, best[0], P * F / best[0], [round(x, 2) for x in best[1]], ["%.12e" % v for v in G[:, 0].tolist()[:3]]))
