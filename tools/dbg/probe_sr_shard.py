"""One rank's share of the strong-scaled cfg3 step (512 CH4 x 1e5 frames over WORLD ranks) on one GPU, per stage.
python tools/dbg/probe_sr_shard.py [world]"""
import sys, time, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth
dev = torch.device("cuda")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nmol, F = 512, 100000
box = synth.methane_box(nmol, seed=3).to(dev)
data = box.trajectory(F + 1, chunk=2048)
mol_atoms = torch.arange(nmol * 5, dtype=torch.int32, device=dev).reshape(nmol, 5)
mass = torch.tensor([12.010635561280681] + [1.0079407573975143] * 4, dtype=torch.float64, device=dev)
axis_atoms = torch.stack([mol_atoms[:, 0], mol_atoms[:, 1], mol_atoms[:, 0], mol_atoms[:, 2]], 1).contiguous()
L = box.celldm
lo, hi = 0, nmol // world
tdev = torch.arange(1, F + 1, device=dev, dtype=torch.float64) * 0.001
M = ops.fft_length(F)
nchunk = max(1, min(8, (hi - lo) // 16))
cuts = [lo + ((hi - lo) * k) // nchunk // 2 * 2 for k in range(nchunk)] + [hi]
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
def step(sync):
    spec = torch.zeros((3, M), dtype=torch.float64, device=dev)
    for k in range(nchunk):
        a, b = cuts[k], cuts[k + 1]
        t0 = time.perf_counter()
        J, _, _ = ops.sr_angmom(data, mol_atoms[a:b].contiguous(), mass, axis_atoms[a:b].contiguous(), 0, (L, L, L),
                                1000.0, want_omega=False, want_axes=False)
        if sync: tick("angmom", t0)
        t0 = time.perf_counter()
        for c in range(3):
            ops.wk_acf_sum(J[:, c, :], pack_real_pairs=True, M=M, spectrum=spec[c])
        if sync: tick("wk", t0)
    t0 = time.perf_counter()
    acf = ops.ifft_acf(spec, F, 1.0 / (float(M) * F * nmol))
    if sync: tick("ifft", t0)
    t0 = time.perf_counter()
    r = ops.simpson(acf, tdev, "cartwright").cpu()
    if sync: tick("simpson", t0)
    return r
for _ in range(3): step(False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step(False)
torch.cuda.synchronize(); print("world %d: %d molecules, %d chunks: %.2f ms per step" % (world, hi - lo, nchunk, (time.perf_counter() - t0) / 5 * 1e3))
for _ in range(5): step(True)
print({k: round(v / 5 * 1e3, 3) for k, v in T.items()})
