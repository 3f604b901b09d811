"""Long trajectories (M1 = 256 / 512): generic word-array route (DYNPY_DD_PATH=v1) against the two-level
thread-per-column kernel (default)."""
import os, sys, time, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 400
F = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
box = synth.water_box(32, seed=2).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
w = ops.wrap(box.trajectory(F, chunk=1 << 16, atom_mask=hidx), box.celldm)
L = box.celldm
i, j = np.triu_indices(64, k=1)
pi = torch.tensor(i[:P], dtype=torch.int32).cuda(); pj = torch.tensor(j[:P], dtype=torch.int32).cuda()
P = pi.numel()
lib = _lib.load()
M = ops.fft_length(F)
res = {}
for mode in ("v1", "v2", "v1", "v2"):
    os.environ["DYNPY_DD_PATH"] = mode
    lib.dynpy_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.time()
    spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M)
    torch.cuda.synchronize(); dt = time.time() - t0
    ms = (ctypes.c_double * 4)(); ln = (ctypes.c_int64 * 4)()
    lib.dynpy_profile_read(ms, ln); lib.dynpy_profile_enable(0)
    res[mode] = ops.ifft_acf(spec, F, 1.0 / (M * F)).cpu().numpy()
    print("M %d path=%s time %.4f s pf/s %.3e  words/cols/rows ms %s" % (M, mode, dt, P * F / dt, [round(x, 2) for x in ms[:3]]))
err = [np.max(np.abs(res["v2"][c] - res["v1"][c])) / np.max(np.abs(res["v1"][c])) for c in range(7)]
print("v2 vs v1", ["%.1e" % e for e in err])
