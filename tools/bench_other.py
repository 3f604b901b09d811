"""Throughput of the quadrupolar (cfg4 shape) and spin-rotation (cfg3 shape) hot paths on one GPU.
Not the headline metric (bench.py measures the dipolar path); numbers go to DESIGN.md."""
import sys, time, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from dynpy_b200 import ops, synth, _lib
lib = _lib.load()
dev = torch.device("cuda")

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, out

def prof(fn):
    lib.dynpy_profile_enable(1); lib.dynpy_profile_read(None, None)
    fn(); ms = (ctypes.c_double * 4)(); ln = (ctypes.c_int64 * 4)()
    lib.dynpy_profile_read(ms, ln); lib.dynpy_profile_enable(0)
    return [round(x, 2) for x in ms[:3]], list(ln[:3])

# ---- quadrupolar: S nuclei x 65536 frames x 6 EFG components (cfg4 shape, S reduced to fit one call)
S, F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 65536
efg = synth.efg_series(256, F, seed=4, device=dev).repeat(S // 256, 1, 1).contiguous()
efg += 1e-3 * torch.randn_like(efg)
ms, spec = timed(lambda: ops.qr_efg_to_spectrum(efg))
print("QR  %d nuclei x %d frames: %.1f ms  %.3e nucleus-frames/s  (48 B each -> %.0f GB/s algorithmic)  cols/rows %s" %
      (S, F, ms, S * F / ms * 1e3, S * F * 48 / ms * 1e3 / 1e9, prof(lambda: ops.qr_efg_to_spectrum(efg))))
del efg, spec
# ---- spin-rotation: 512 CH4 x 1e5 frames (cfg3 shape): angular momenta + 3 ACFs per molecule
nmol, F = 512, 100001
box = synth.methane_box(nmol, seed=3).to(dev)
pos = box.trajectory(F, chunk=2048)
mol_atoms = torch.arange(nmol * 5, dtype=torch.int32, device=dev).reshape(nmol, 5)
mass = torch.tensor([12.010635561280681] + [1.0079407573975143] * 4, dtype=torch.float64, device=dev)
axis_atoms = torch.stack([mol_atoms[:, 0], mol_atoms[:, 1], mol_atoms[:, 0], mol_atoms[:, 2]], 1).contiguous()
L = box.celldm
def sr():
    J, om, ax = ops.sr_angmom(pos, mol_atoms, mass, axis_atoms, 0, (L, L, L), 1000.0, want_omega=False, want_axes=False)
    M = ops.fft_length(F - 1)
    spec = torch.zeros((3, M), dtype=torch.float64, device=dev)
    for c in range(3):
        ops.wk_acf_sum(J[:, c, :].contiguous(), pack_real_pairs=True, M=M, spectrum=spec[c])
    return ops.ifft_acf(spec, F - 1, 1.0 / (float(M) * (F - 1) * nmol))
ms_k, _ = timed(lambda: ops.sr_angmom(pos, mol_atoms, mass, axis_atoms, 0, (L, L, L), 1000.0, want_omega=False, want_axes=False))
print("SR  k_sr_angmom alone: %.2f ms (%.0f GB/s of the 120 B per molecule-frame)" % (ms_k, nmol * (F - 1) * 120 / ms_k * 1e3 / 1e9))
ms, acf = timed(sr)
print("SR  %d CH4 x %d frames: %.1f ms  %.3e molecule-frames/s  (120 B each -> %.0f GB/s algorithmic)  cols/rows %s" %
      (nmol, F - 1, ms, nmol * (F - 1) / ms * 1e3, nmol * (F - 1) * 120 / ms * 1e3 / 1e9, prof(sr)))
