import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 300
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
box = synth.water_box(32, seed=2).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
raw = box.trajectory(F, chunk=4096, atom_mask=hidx)
L = box.celldm
w = ops.wrap(raw, L)
i, j = np.triu_indices(64, k=1)
pi = torch.tensor(i[:P], dtype=torch.int32).cuda(); pj = torch.tensor(j[:P], dtype=torch.int32).cuda()
M = ops.fft_length(F)
res = {}
for mode, env in (("0", {"DYNPY_DD_STREAM": "0", "DYNPY_DD_FUSED_COLS": "0"}), ("fused", {"DYNPY_DD_STREAM": "0", "DYNPY_DD_FUSED_COLS": "1"}), ("1", {"DYNPY_DD_STREAM": "1"})):
    os.environ.update(env)
    torch.cuda.synchronize(); t0 = time.time()
    spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M)
    torch.cuda.synchronize(); dt = time.time() - t0
    G = ops.ifft_acf(spec, F, 1.0 / (M * F)).cpu().numpy()
    res[mode] = G
    print("mode", mode, "time %.3f s" % dt, "pf/s %.3e" % (P * F / dt))
for m in ("fused", "1"):
    err = [np.max(np.abs(res[m][c] - res["0"][c])) / np.max(np.abs(res["0"][c])) for c in range(7)]
    print(m, "vs two-kernel normwise err", ["%.1e" % e for e in err])
