"""Throughput of the dipolar fast path as a function of the T-workspace budget (batch size)."""
import os, sys, time, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
box = synth.water_box(32, seed=2).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
raw = box.trajectory(F, chunk=4096, atom_mask=hidx)
L = box.celldm
w = ops.wrap(raw, L)
i, j = np.triu_indices(64, k=1)
pi = torch.tensor(i[:P], dtype=torch.int32).cuda(); pj = torch.tensor(j[:P], dtype=torch.int32).cuda()
P = pi.numel()
M = ops.fft_length(F)
lib = _lib.load()
for mb in [int(a) for a in sys.argv[3:]] or [64, 128, 256, 512, 1024, 2048, 4096, 8192]:
    for rep in range(2):
        lib.dynpy_profile_enable(1)
        torch.cuda.synchronize(); t0 = time.time()
        spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M, ws_budget=mb << 20)
        torch.cuda.synchronize(); dt = time.time() - t0
        ms = (ctypes.c_double * 4)(); ln = (ctypes.c_int64 * 4)()
        lib.dynpy_profile_read(ms, ln)
        lib.dynpy_profile_enable(0)
    print("budget %5d MB  time %.4f s  pf/s %.3e  cols/rows ms %s launches %s" % (mb, dt, P * F / dt, [round(x, 2) for x in ms[1:3]], list(ln[1:3])))
