#!/usr/bin/env python
"""cuFFT yardstick for the hand-written transform kernels — a TOOL, not the product path.

Same box, same run, fp64, CUDA-event timing after warm-up:

  (a) raw transforms: torch.fft.fft (cuFFT Z2Z) of B zero-padded complex series, M = 204800 (cfg2) and 2^21
      (cfg5), next to dynpy_wk_acf_sum_f64 (forward transform + |X|^2 + sum over series in one call) on the
      same series -> ps per transform point for both;
  (b) the dipolar step as one would write it with library calls: pair vectors / Y2m r^-3 words by torch
      elementwise ops, torch.fft.fft / rfft (n = M), |X|^2 summed over pairs — next to
      dynpy_dd_pairs_to_spectrum_f64 on the same pairs -> pair-frames/s for both.

    python tools/cufft_yardstick.py > profiles/r02_cufft_yardstick.json
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dynpy_b200 import ops, synth  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def raw_transforms(N, B):
    """B complex series of N samples, zero-padded to M: cuFFT vs the hand-written accumulate-power call."""
    M = ops.fft_length(N)
    g = torch.Generator(device="cuda").manual_seed(1)
    re = torch.randn((B, N), dtype=torch.float64, device="cuda", generator=g)
    im = torch.randn((B, N), dtype=torch.float64, device="cuda", generator=g)
    z = torch.complex(re, im)
    zp = torch.zeros((B, M), dtype=torch.complex128, device="cuda")
    zp[:, :N] = z
    out = {"N": N, "M": M, "B": B}
    # cuFFT alone on an already padded buffer
    ms = timed(lambda: torch.fft.fft(zp, dim=-1))
    out["cufft_only_ps_per_point"] = ms * 1e9 / (B * M)
    # cuFFT with the padding (n = M) and the |X|^2 + sum passes the ACF needs

    def lib():
        X = torch.fft.fft(z, n=M, dim=-1)
        return (X.real * X.real + X.imag * X.imag).sum(0)
    ms = timed(lib)
    out["cufft_pad_power_sum_ps_per_point"] = ms * 1e9 / (B * M)
    spec = torch.zeros(M, dtype=torch.float64, device="cuda")
    ms = timed(lambda: ops.wk_acf_sum(re, im, M=M, spectrum=spec))
    out["hand_wk_acf_sum_ps_per_point"] = ms * 1e9 / (B * M)
    # same answer?
    ref = lib()
    spec.zero_()
    ops.wk_acf_sum(re, im, M=M, spectrum=spec)
    k = torch.arange(M, device="cuda")
    M1 = M // 4096 if M > 4096 else 1
    nat = spec.reshape(M1, -1).t().reshape(-1) if M > 4096 else spec  # internal order k1*4096+k2 <-> k = k1 + M1 k2
    out["max_rel_diff_vs_cufft"] = float((nat - ref).abs().max() / ref.abs().max())
    del k
    return out


def torch_dd_spectrum(w, pi, pj, L, M, chunk):
    """Library-call restatement of the fused dipolar call: returns ([3, M] complex-series power, [4, M/2+1]
    real-series power)."""
    F = w.shape[2]
    Sc = torch.zeros((4, M), dtype=torch.float64, device=w.device)
    Sr = torch.zeros((3, M // 2 + 1), dtype=torch.float64, device=w.device)
    c20, c21, c22 = 0.25 * np.sqrt(5 / np.pi), -0.5 * np.sqrt(15 / (2 * np.pi)), 0.25 * np.sqrt(15 / (2 * np.pi))
    kf = np.sqrt(4 * np.pi / 5)
    for s0 in range(0, pi.numel(), chunk):
        a, b = w[pi[s0:s0 + chunk].long()], w[pj[s0:s0 + chunk].long()]
        d = a - b
        d -= L * torch.round(d / L)
        r2 = (d * d).sum(1)
        rinv = torch.rsqrt(r2)
        x, y, z = d[:, 0] * rinv, d[:, 1] * rinv, d[:, 2] * rinv
        r3 = rinv ** 3
        y20 = c20 * (3 * z * z - 1)
        xy = torch.complex(x, y)
        y21 = c21 * z * xy
        y22 = c22 * xy * xy
        for k, s in enumerate((y21, y22, kf * r3 * y21, kf * r3 * y22)):
            X = torch.fft.fft(s, n=M, dim=-1)
            Sc[k] += (X.real * X.real + X.imag * X.imag).sum(0)
        for k, s in enumerate((y20, r3 - r3.mean(1, keepdim=True), kf * r3 * y20)):
            X = torch.fft.rfft(s, n=M, dim=-1)
            Sr[k] += (X.real * X.real + X.imag * X.imag).sum(0)
    return Sc, Sr


def dipolar(nmol, F, P, chunk):
    box = synth.water_box(nmol, seed=2).to("cuda")
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device="cuda")
    L = box.celldm
    w = ops.wrap(box.trajectory(F, chunk=1 << 14, atom_mask=hidx), L)
    i, j = np.triu_indices(2 * nmol, k=1)
    pi = torch.from_numpy(i[:P].astype(np.int32)).cuda()
    pj = torch.from_numpy(j[:P].astype(np.int32)).cuda()
    M = ops.fft_length(F)
    out = {"frames": F, "M": M, "pairs": P, "torch_chunk_pairs": chunk}
    ms = timed(lambda: torch_dd_spectrum(w, pi, pj, L, M, chunk), reps=2)
    out["cufft_path_pair_frames_per_s"] = P * F / (ms * 1e-3)
    ms = timed(lambda: ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M), reps=2)
    out["hand_path_pair_frames_per_s"] = P * F / (ms * 1e-3)
    out["hand_over_cufft"] = out["hand_path_pair_frames_per_s"] / out["cufft_path_pair_frames_per_s"]
    # agreement of the two on the complex Y21 power spectrum (natural order vs internal order)
    Sc, Sr = torch_dd_spectrum(w, pi, pj, L, M, chunk)
    spec = ops.dd_pairs_to_spectrum(w, pi, pj, (L, L, L), M=M)
    M1 = M // 4096
    nat = spec[1].reshape(M1, -1).t().reshape(-1)
    out["max_rel_diff_Y21_power"] = float((nat - Sc[0]).abs().max() / Sc[0].abs().max())
    return out


def main():
    torch.cuda.set_device(0)
    res = {"tool": "cufft_yardstick", "gpu": torch.cuda.get_device_name(0), "dtype": "f64 / complex128",
           "raw": [raw_transforms(100_000, 512), raw_transforms(1_000_000, 48)],
           "dipolar": [dipolar(32, 100_000, 1024, 32), dipolar(8, 1_000_000, 96, 4)]}
    for r in res["raw"]:
        r["hand_over_cufft_pad_power_sum"] = r["cufft_pad_power_sum_ps_per_point"] / r["hand_wk_acf_sum_ps_per_point"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
