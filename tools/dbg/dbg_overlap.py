"""Does running the dipolar call on S concurrent streams (disjoint pair halves, own workspaces) beat one stream?"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
F = 100000
box = synth.water_box(32, seed=2).to("cuda")
hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
w = ops.wrap(box.trajectory(F, chunk=4096, atom_mask=hidx), box.celldm)
L = box.celldm
i, j = np.triu_indices(64, k=1)
pi = torch.tensor(i[:P], dtype=torch.int32).cuda(); pj = torch.tensor(j[:P], dtype=torch.int32).cuda()
P = pi.numel()
lib = _lib.load()
M = ops.fft_length(F)
A = w.shape[0]


def run(nstreams, items):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    wsb = int(lib.dynpy_dd_workspace_bytes(F, M, items))
    wss = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(nstreams)]
    specs = [torch.zeros((7, M), dtype=torch.float64, device="cuda") for _ in range(nstreams)]
    cuts = np.linspace(0, P, nstreams + 1).astype(int)
    cuts = (cuts // 2) * 2
    cuts[-1] = P
    best = 1e9
    for rep in range(3):
        for s in specs:
            s.zero_()
        torch.cuda.synchronize(); t0 = time.time()
        # interleave the submissions batch by batch so that the streams really run side by side
        step = 2 * items
        pos = [int(c) for c in cuts[:-1]]
        live = True
        while live:
            live = False
            for k, st in enumerate(streams):
                if pos[k] < cuts[k + 1]:
                    n = min(step, int(cuts[k + 1]) - pos[k])
                    rc = lib.dynpy_dd_pairs_to_spectrum_f64(w.data_ptr(), A, F, pi[pos[k]:].data_ptr(), pj[pos[k]:].data_ptr(), n,
                                                            L, L, L, specs[k].data_ptr(), M, wss[k].data_ptr(), wsb, st.cuda_stream)
                    assert rc == 0
                    pos[k] += n
                    live = True
        torch.cuda.synchronize(); best = min(best, time.time() - t0)
    return best, sum(specs)


t1, s1 = run(1, 100)
print("1 stream  x100 items: %.4f s  %.3e pf/s" % (t1, P * F / t1))
for ns, items in ((2, 50), (2, 100), (3, 33)):
    t, s = run(ns, items)
    err = float((s - s1).abs().max() / s1.abs().max())
    print("%d streams x%d items: %.4f s  %.3e pf/s  (%.1f%%)  err %.1e" % (ns, items, t, P * F / t, 100 * (t1 / t - 1), err))
