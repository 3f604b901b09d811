"""A/B check of the dipolar fast path (dd2) against the previous generation kernels on the GPU."""
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
res = {}
modes = (("v1", {"DYNPY_DD_PATH": "v1"}), ("v2", {"DYNPY_DD_PATH": "v2"}))
if len(sys.argv) > 3:
    modes = tuple(m for m in modes if m[0] == sys.argv[3])
for mode, env in modes:
    os.environ.update(env)
    for rep in range(2):
        lib.dynpy_profile_enable(1)
        torch.cuda.synchronize(); t0 = time.time()
        spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M)
        torch.cuda.synchronize(); dt = time.time() - t0
        ms = (ctypes.c_double * 4)(); ln = (ctypes.c_int64 * 4)()
        lib.dynpy_profile_read(ms, ln)
        lib.dynpy_profile_enable(0)
    G = ops.ifft_acf(spec, F, 1.0 / (M * F)).cpu().numpy()
    res[mode] = G
    print("mode", mode, "time %.4f s" % dt, "pf/s %.3e" % (P * F / dt), "words/cols/rows ms", [round(x, 2) for x in ms[:3]], list(ln[:3]))
if F > 65536 and F <= 102400:
    os.environ["DYNPY_DD_PATH"] = "v2"
    M2 = 1 << 18
    spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M2)
    Gp = ops.ifft_acf(spec, F, 1.0 / (M2 * F)).cpu().numpy()
    for m in res:
        err = [np.max(np.abs(res[m][c] - Gp[c])) / np.max(np.abs(Gp[c])) for c in range(7)]
        print(m, "(M=%d) vs v2 at M=2^18, normwise err" % M, ["%.1e" % e for e in err])
if len(res) < 2:
    sys.exit(0)
err = [np.max(np.abs(res["v2"][c] - res["v1"][c])) / np.max(np.abs(res["v1"][c])) for c in range(7)]
print("v2 vs v1 normwise err", ["%.1e" % e for e in err])
