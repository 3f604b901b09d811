"""Per-frame cost of the nearest-molecule cluster extraction at a realistic size (256 waters + 1 ion)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import neighbors, parseMD, synth, ops
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = 4
box = synth.water_box(nmol, seed=2)
L = box.celldm
raw = box.frames(0, F)
g = np.stack(np.meshgrid(*(np.linspace(0, L, 24, endpoint=False),) * 3, indexing="ij"), -1).reshape(-1, 3)
at = raw[:, :, 0].numpy()
d = g[:, None, :] - at[None, :, :]
d -= L * np.round(d / L)
hole = g[np.argmax(np.sqrt((d ** 2).sum(-1)).min(1))]
ion = torch.tensor(hole, dtype=torch.float64)[None, :, None].expand(1, 3, F)
raw = torch.cat([raw, ion], 0)
tr = parseMD.Trajectory(pos=raw.cuda().contiguous(), symbols=box.symbols + ["Na"], frames=np.arange(F, dtype=np.int64))
A = tr.nat
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    dct = neighbors.periodic_nearest_neighbors_by_atom(tr, [A - 1], L, [30], dmax=L / 2, Na=0.1)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("A=%d (27A=%d): %.3f s per frame, nearest rows %d, cluster atoms per frame %d, peak mem %.2f GB"
          % (A, 27 * A, dt / F, len(dct["nearest"]), len(dct[30].atom) // F, torch.cuda.max_memory_allocated() / 2**30))
# device part only
x27 = ops.supercell(tr.pos, 0, L); u27 = ops.wrap(x27, 3 * L)
idx = torch.arange(27 * A, dtype=torch.int64, device="cuda")
torch.cuda.synchronize(); t0 = time.time()
out = ops.pdist_ortho(u27[0], u27[1], u27[2], 3 * L, 3 * L, 3 * L, idx, L / 2)
torch.cuda.synchronize(); print("pair table of the image block: %.3f s, %d rows of %d candidates" % (time.time() - t0, out[0].numel(), 27 * A * (27 * A - 1) // 2))
