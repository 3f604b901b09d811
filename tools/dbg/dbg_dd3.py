"""A/B of the two dipolar column kernels (DYNPY_DD_COLS=warp | tpc) on the GPU."""
import os, sys, time, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
only = sys.argv[3] if len(sys.argv) > 3 else None
box = synth.water_box(32, seed=2).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
raw = box.trajectory(F, chunk=4096, atom_mask=hidx)
L = box.celldm
w = ops.wrap(raw, L)
i, j = np.triu_indices(64, k=1)
pi = torch.tensor(i[:P], dtype=torch.int32).cuda(); pj = torch.tensor(j[:P], dtype=torch.int32).cuda()
P = pi.numel()
lib = _lib.load()
res = {}
Ms = [ops.fft_length(F)] + ([1 << 18] if 65536 < F <= 102400 else [])
for M in Ms:
    for mode in ("warp", "tpc"):
        if only and mode != only:
            continue
        os.environ["DYNPY_DD_COLS"] = mode
        for rep in range(2):
            lib.dynpy_profile_enable(1)
            torch.cuda.synchronize(); t0 = time.time()
            spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M)
            torch.cuda.synchronize(); dt = time.time() - t0
            ms = (ctypes.c_double * 4)(); ln = (ctypes.c_int64 * 4)()
            lib.dynpy_profile_read(ms, ln)
            lib.dynpy_profile_enable(0)
        res[(M, mode)] = ops.ifft_acf(spec, F, 1.0 / (M * F)).cpu().numpy()
        print("M %d cols=%s time %.4f s pf/s %.3e  geom/cols/rows ms %s %s" % (M, mode, dt, P * F / dt, [round(x, 2) for x in ms[:3]], list(ln[:3])))
keys = list(res)
for k in keys[1:]:
    err = [np.max(np.abs(res[k][c] - res[keys[0]][c])) / np.max(np.abs(res[keys[0]][c])) for c in range(7)]
    print(k, "vs", keys[0], ["%.1e" % e for e in err])
