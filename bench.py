#!/usr/bin/env python
"""Benchmark of the dipolar G(tau)+R1 hot path (BASELINE.json metric: pair-frames/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" = one pass of the hot path over the whole workload: wrapped coordinates resident in
HBM -> pair vectors -> Y2m r^-3 words -> batched FFTs -> accumulated spectra -> summed ACFs
G(tau) (7 columns) -> Simpson J(0) -> 1/T1.  Under torchrun every rank holds the (replicated)
coordinates, transforms a contiguous shard of the pairs and the spectra-derived ACFs are
all-reduced once (NCCL); the timed region is bracketed by barrier + synchronize and the
reported time is the max over ranks.

Workloads: cfg2 (default) 256 H2O x 1e5 frames, all 130 816 H-H pairs, gapless
rmax; cfg5 512 H2O x 1e6 frames, 523 264 intermolecular H-H pairs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_PER_PAIR_FRAME = 48.0  # two fp64 xyz triplets per pair-frame

WORKLOADS = {
    "cfg2": dict(nmol=256, frames=100_000, pair_type="all", seed=2,
                 name="cfg2: dipolar, 256 H2O x 1e5 frames, all H-H pairs, gapless rmax"),
    "cfg5": dict(nmol=512, frames=1_000_000, pair_type="inter", seed=5,
                 name="cfg5: dipolar, 512 H2O x 1e6 frames, intermolecular H-H pairs, gapless rmax"),
    "small": dict(nmol=32, frames=20_000, pair_type="all", seed=2,
                  name="small: dipolar, 32 H2O x 2e4 frames, all H-H pairs, gapless rmax"),
    # the other two mechanisms (not the headline metric; same JSON contract, their own units)
    "cfg3": dict(kind="sr", nmol=512, frames=100_000, seed=3,
                 name="cfg3: spin-rotation, 512 CH4 x 1e5 frames (angular momenta + 3 ACFs per molecule + rate)"),
    "cfg4": dict(kind="qr", nuclei=12_500, frames=65_536, seed=4,
                 name="cfg4: quadrupolar, 1e5 nuclei over 8 GPUs = 12500 nuclei per GPU x 65536 frames x 6 EFG components "
                      "(5 ACFs + rates)"),
}


def h_pairs(nmol, pair_type):
    """H-only atom indexing: H atoms 2m, 2m+1 belong to molecule m; pairs i<j."""
    nh = 2 * nmol
    i, j = np.triu_indices(nh, k=1)
    if pair_type == "inter":
        keep = (i // 2) != (j // 2)
        i, j = i[keep], j[keep]
    elif pair_type == "intra":
        keep = (i // 2) == (j // 2)
        i, j = i[keep], j[keep]
    return i.astype(np.int32), j.astype(np.int32)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    def __init__(self, index):
        self.path = "/tmp/dynpy_clocks_%d.csv" % os.getpid()
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap,power.draw"
        try:
            self.fh = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.fh,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                t = [x.strip() for x in ln.split(",")]
                if len(t) < 7:
                    continue
                sm.append(float(t[0])); mx.append(float(t[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------ reference arm
_REF_SHARED = {}


def _ref_worker(pairs):
    # the trajectory is inherited through fork (no per-task pickling of the coordinates)
    from oracle import oracle_np as orc
    return orc.dd_pairs_gapless(_REF_SHARED["w"], pairs, _REF_SHARED["L"])


def reference_arm(args, wl):
    """The reference's algorithm (oracle port: per-pair min-image vectors, angle-form Y2m, one
    zero-padded fft/ifft per series as in wiener_khinchin) on the host cores, one process per core
    over disjoint pair shards, on a bounded sample of the workload's pairs."""
    import multiprocessing as mp

    import torch

    from dynpy_b200 import synth
    from oracle import oracle_np as orc  # noqa: F401
    cores = os.cpu_count() or 1
    F = wl["frames"] if wl["frames"] <= 100_000 else 100_000
    nh_sample = 32
    box = synth.water_box(wl["nmol"], seed=wl["seed"])
    hmask = torch.tensor([s == "H" for s in box.symbols])
    traj = box.trajectory(F, atom_mask=hmask.nonzero()[:, 0][:nh_sample]).numpy()
    L = box.celldm
    w = np.mod(traj, L)
    pi, pj = h_pairs(nh_sample // 2, wl["pair_type"])
    per_step = min(len(pi), max(cores * 8, 8))   # 8 pairs per core and step: a few seconds of work per step
    shards = [list(zip(pi[k:per_step:cores].tolist(), pj[k:per_step:cores].tolist())) for k in range(cores)]
    shards = [s for s in shards if s]
    _REF_SHARED.update(w=w, L=L)
    times = []
    with mp.get_context("fork").Pool(len(shards)) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            parts = pool.map(_ref_worker, shards)
            G = sum(parts)
            t = np.arange(F) * 0.001
            orc.dd_rates(t, G[4] * 2 / nh_sample, G[5] * 2 / nh_sample, G[6] * 2 / nh_sample)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = per_step * F / (ms * 1e-3)
    sample = "%d pairs x %d frames per step (of %s)" % (per_step, F, wl["name"])
    line = {"impl": "reference", "metric": "pair-frames/s (dipolar G(tau)+R1)", "value": value, "unit": "pair-frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "pair-frames/s", "cores": len(shards), "kind": "port",
                             "sample": sample, "reference_shim_recorded": shim_record("dd")},
            "e2e": {"value": value, "unit": "pair-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def shim_record(kind):
    """What oracle/time_reference.py measured in the build container (the GPU box has no /root/reference):
    the unmodified reference through the pandas shim and the numpy port on the same inputs."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_cpu_reference_shim.json")) as fh:
            c = json.load(fh)["cases"][kind]
        return {"case": c["case"], "reference_units_per_s": c["reference_shim_units_per_s"], "reference_cores": c["reference_cores"],
                "port_units_per_s_same_inputs": c["port_units_per_s"], "port_over_reference": c["port_over_reference"],
                "where": "build container, 8 vCPU (profiles/r02_cpu_reference_shim.json)"}
    except Exception:
        return None


def _ref_qr_worker(job):
    """Quadrupolar port on a block of nuclei: R_2m from the EFG components, five zero-padded fft/ifft ACFs each
    (qrax.py:115-166), summed."""
    from oracle import oracle_np as orc
    efg = _REF_SHARED["efg"]
    lo, hi = job
    acc = None
    for n in range(lo, hi):
        V = {k: efg[n, c] for c, k in enumerate(("Vxx", "Vyy", "Vzz", "Vxy", "Vxz", "Vyz"))}
        ac = np.stack([orc.wiener_khinchin(r) for r in orc.qr_spatial(V)])
        acc = ac if acc is None else acc + ac
    return acc


def _ref_sr_worker(job):
    """Spin-rotation port on a small methane box of its own (sr_pipeline: pair table, molecules, SR_func1 per
    molecule-frame, three ACFs per molecule, mean, rate; SRparse.py:25-262)."""
    from dynpy_b200 import synth
    from oracle import oracle_np as orc
    seed, nmol, F = job
    box = synth.methane_box(nmol, seed=seed)
    pos = box.frames(0, F + 1).numpy().transpose(2, 0, 1)
    out = orc.sr_pipeline(pos, box.symbols, np.arange(1, F + 2), box.celldm, 0.001, "methane", nmol,
                          [16.495, 16.495, -1.875])
    return out["rate"]


def reference_arm_other(args, wl):
    """--impl reference for the quadrupolar / spin-rotation workloads: the numpy port on all host cores, each step
    a bounded sample of the workload's shape (cfg4: cores x 16 nuclei x 65536 frames; cfg3: cores x 8 CH4 x 500
    frames — the port's pair table is O(A^2) per frame, so the sample keeps the box small)."""
    import multiprocessing as mp

    from dynpy_b200 import synth
    from oracle import oracle_np as orc
    cores = os.cpu_count() or 1
    if wl["kind"] == "qr":
        F = wl["frames"]
        per_core = 16
        S = cores * per_core
        _REF_SHARED["efg"] = synth.efg_series(S, F, seed=wl["seed"]).numpy()
        jobs = [(k * per_core, (k + 1) * per_core) for k in range(cores)]
        worker, units, unit = _ref_qr_worker, float(S) * F, "nucleus-frames/s"
        metric = "nucleus-frames/s (quadrupolar ACFs + rates)"
        sample = "%d nuclei x %d frames per step (of %s)" % (S, F, wl["name"])
        kind = "qr"
    else:
        nmol, F = 8, 500
        jobs = [(wl["seed"] + k, nmol, F) for k in range(cores)]
        worker, units, unit = _ref_sr_worker, float(cores * nmol) * F, "molecule-frames/s"
        metric = "molecule-frames/s (spin-rotation ACFs + rate)"
        sample = "%d boxes of %d CH4 x %d frames per step (of %s)" % (cores, nmol, F, wl["name"])
        kind = "sr"
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            parts = pool.map(worker, jobs)
            if kind == "qr":
                f = sum(parts) / S
                orc.qr_from_acf(np.arange(F) * 0.029, f, "127I")
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = units / (ms * 1e-3)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if kind == "qr" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                             "reference_shim_recorded": shim_record(kind)},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline_sample(box, wl, F, budget_s=20.0):
    """Single-core oracle port on a bounded sample (a few pairs x all frames)."""
    import torch

    from oracle import oracle_np as orc
    F = min(F, 100_000)
    nh = 8
    hmask = torch.tensor([s == "H" for s in box.symbols])
    traj = box.trajectory(F, atom_mask=hmask.nonzero()[:, 0][:nh].to(box.lattice.device)).cpu().numpy()
    w = np.mod(traj, box.celldm)
    pi, pj = h_pairs(nh // 2, "all")
    pairs = list(zip(pi.tolist(), pj.tolist()))
    t0 = time.perf_counter()
    done = 0
    for p in pairs:
        orc.dd_pairs_gapless(w, [p], box.celldm)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done * F / dt, "unit": "pair-frames/s", "cores": 1, "kind": "port",
            "sample": "%d pairs x %d frames of the workload, numpy restatement of the reference path, 1 core" % (done, F),
            "reference_shim_recorded": shim_record("dd")}


def check_result(ops, dist, wrapped, pi, pj, cell, F, M, G_all, nspins, lo, hi, rank):
    """Parity gate of the bench line (raises instead of printing a number for a wrong result):
    (1) the fused CUDA call on a subset of the workload's pairs (first, middle, last) against the CPU oracle —
        full ACFs of all 7 columns, normwise <= 1e-9, and the rate from them <= 1e-8;
    (2) lag 0 of the WHOLE job's seven summed ACFs against sum_pairs mean_t |series|^2 formed by independent
        torch code over ALL pairs (Parseval: size-independent, covers every pair of the workload)."""
    import torch

    from oracle import oracle_np as orc
    L = cell[0]
    P = pi.numel()
    out = {}
    if rank == 0:
        out = _check_subset_vs_oracle(ops, orc, wrapped, pi, pj, cell, F, M)
    # ---- Parseval over all pairs (every rank sums its own shard)
    tot = torch.zeros(7, dtype=torch.float64, device=wrapped.device)
    kf = 4 * np.pi / 5
    c20, c21, c22 = (0.25 * np.sqrt(5 / np.pi)) ** 2, (0.5 * np.sqrt(15 / (2 * np.pi))) ** 2, (0.25 * np.sqrt(15 / (2 * np.pi))) ** 2
    chunk = max(1, int(2.5e7 // F))
    for s0 in range(lo, hi, chunk):
        s1 = min(s0 + chunk, hi)
        d = wrapped[pi[s0:s1].long()] - wrapped[pj[s0:s1].long()]
        d -= L * torch.round(d / L)
        r2 = (d * d).sum(dim=1)
        r3 = r2 ** -1.5
        z2 = d[:, 2] ** 2 / r2
        rho2 = 1.0 - z2
        y = torch.stack([c20 * (3 * z2 - 1) ** 2, c21 * z2 * rho2, c22 * rho2 * rho2])
        dr3 = r3 - r3.mean(dim=1, keepdim=True)
        tot[0:3] += y.mean(dim=2).sum(dim=1)
        tot[3] += (dr3 * dr3).mean(dim=1).sum()
        tot[4:7] += (kf * y * (r3 * r3)[None]).mean(dim=2).sum(dim=1)
    dist.allreduce_sum_(tot)
    want = (tot * (2.0 / nspins)).cpu().numpy()
    got = G_all[:, 0].cpu().numpy()
    den = np.abs(want)
    den[3] = max(den[3], 1e-6 * den[4:7].max())   # G_r of rigid pairs is rounding noise: judge it on the F2m scale
    rel = np.abs(got - want) / den
    if float(rel.max()) > 1e-9:
        raise RuntimeError("bench parity check failed: G(0) %r vs direct sums %r" % (got, want))
    out.update(parseval_pairs=int(P), parseval_max_rel_err=float(rel.max()))
    return out


def _check_subset_vs_oracle(ops, orc, wrapped, pi, pj, cell, F, M):
    import torch
    L = cell[0]
    P = pi.numel()
    sub = sorted({0, 1, P // 2, P - 1})
    spi, spj = pi[sub].contiguous(), pj[sub].contiguous()
    spec = ops.dd_pairs_to_spectrum(wrapped, spi, spj, cell, M=M)
    Gs = ops.ifft_acf(spec, F, 1.0 / (float(M) * F)).cpu().numpy()
    w_sub_atoms = torch.unique(torch.cat([spi, spj])).long()
    remap = {int(a): k for k, a in enumerate(w_sub_atoms.tolist())}
    wh = wrapped[w_sub_atoms].cpu().numpy()
    ref = orc.dd_pairs_gapless(wh, [(remap[int(a)], remap[int(b)]) for a, b in zip(spi.tolist(), spj.tolist())], L)
    scale = max(np.max(np.abs(ref[c])) for c in (4, 5, 6))
    errs = []
    for c in range(7):
        denom = np.max(np.abs(ref[c])) if c != 3 else max(np.max(np.abs(ref[3])), 1e-6 * scale)
        errs.append(float(np.max(np.abs(Gs[c] - ref[c])) / denom))
    if max(errs) > 1e-9:
        raise RuntimeError("bench parity check failed: ACF columns differ from the oracle by %r" % (errs,))
    t = np.arange(1, F + 1) * 0.001
    r_gpu = orc.dd_rates(t, Gs[4], Gs[5], Gs[6])[0]
    r_ref = orc.dd_rates(t, ref[4], ref[5], ref[6])[0]
    if abs(r_gpu - r_ref) > 1e-8 * abs(r_ref):
        raise RuntimeError("bench parity check failed: rate %r vs oracle %r" % (r_gpu, r_ref))
    return {"oracle_pairs": len(sub), "acf_max_normwise_err": max(errs), "rate_rel_err": abs(r_gpu - r_ref) / abs(r_ref)}


def check_ragged(ops, wrapped, pi, pj, cell, rmax, F):
    """Ragged-path gate: a handful of pairs through the CUDA ragged call against the oracle's compressed-series
    restatement (full [7, F] scatter), normwise <= 1e-9."""
    import torch

    from oracle import oracle_np as orc
    P = pi.numel()
    counts = ops.dd_pair_count(wrapped, pi, pj, cell, rmax)
    part = torch.nonzero((counts > 0) & (counts < F))[:, 0]
    if part.numel() == 0:
        return {"ragged_pairs_checked": 0}
    sub = part[torch.linspace(0, part.numel() - 1, 4).long()].unique()
    spi, spj = pi[sub].contiguous(), pj[sub].contiguous()
    got = ops.dd_pairs_ragged_to_acf(wrapped, spi, spj, cell, rmax).cpu().numpy()
    atoms = torch.unique(torch.cat([spi, spj])).long()
    remap = {int(a): k for k, a in enumerate(atoms.tolist())}
    ref = orc.dd_pairs_ragged(wrapped[atoms].cpu().numpy(),
                              [(remap[int(a)], remap[int(b)]) for a, b in zip(spi.tolist(), spj.tolist())], cell[0], rmax)
    scale = max(np.max(np.abs(ref[c])) for c in (4, 5, 6))
    errs = []
    for c in range(7):
        denom = np.max(np.abs(ref[c])) if c != 3 else max(np.max(np.abs(ref[3])), 1e-6 * scale)
        errs.append(float(np.max(np.abs(got[c] - ref[c])) / denom))
    if max(errs) > 1e-9:
        raise RuntimeError("bench parity check failed (ragged path): %r" % (errs,))
    return {"ragged_pairs_checked": int(sub.numel()), "acf_max_normwise_err": max(errs)}


# ------------------------------------------------------------------------------------------ native arm
def native_arm(args, wl):
    import torch
    import torch.distributed as tdist

    from dynpy_b200 import _lib, ddrax, dist, ops, synth
    # stdout carries exactly one JSON line: anything a library prints while the job runs (NCCL's version
    # banner, for one) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    F = wl["frames"] if args.frames is None else args.frames
    box = synth.water_box(wl["nmol"], seed=wl["seed"]).to(dev)
    hidx = torch.tensor([k for k, s in enumerate(box.symbols) if s == "H"], device=dev)
    raw = box.trajectory(F, chunk=2048, atom_mask=hidx)          # [2*nmol, 3, F] raw bohr, resident
    L = box.celldm
    cell = (L, L, L)
    rmax = L if args.rmax is None else float(args.rmax)            # L > sqrt(3)/2 L: every pair in every frame
    ragged = not ddrax.gapless(cell, rmax)
    pi_h, pj_h = h_pairs(wl["nmol"], wl["pair_type"])
    if args.pairs:
        pi_h, pj_h = pi_h[:args.pairs], pj_h[:args.pairs]
    P = len(pi_h)
    pi, pj = torch.from_numpy(pi_h).to(dev), torch.from_numpy(pj_h).to(dev)
    wrapped = torch.empty_like(raw)
    tdev = torch.arange(1, F + 1, device=dev, dtype=torch.float64) * 0.001
    M = ops.fft_length(F)
    lo, hi = dist.shard_range(P, rank, world)
    ws_budget = int(args.ws_gb * (1 << 30)) if args.ws_gb else None     # None: the library's adaptive default

    present = [float(P) * float(F)]

    def step_ragged(src):
        # rmax inside the box: the product's own driver call (presence count, always-present pairs through the
        # fast path, the others through the ragged path, reference quirk B.3), pairs sharded over the ranks
        ops.wrap(src, L, out=wrapped)
        G, hit, nspins, info = ddrax.pair_acf_sum(wrapped, pi, pj, cell, rmax)
        Gk = (G[:, hit] * (2.0 / nspins)).contiguous()
        EN = ddrax.spec_dens_extr_narr(tdev[hit].contiguous(), Gk, "cartwright")
        return Gk, ddrax.dipolar(EN, "1H", "1H")

    def step(src):
        if ragged:
            return step_ragged(src)
        ops.wrap(src, L, out=wrapped)
        # pair_acf_sum with an explicit workspace budget
        G = torch.zeros((7, F), dtype=torch.float64, device=dev)
        if hi > lo:
            spec = ops.dd_pairs_to_spectrum(wrapped, pi[lo:hi], pj[lo:hi], cell, M=M, ws_budget=ws_budget)
            G += ops.ifft_acf(spec, F, 1.0 / (float(M) * float(F)))
        dist.allreduce_sum_(G)
        G *= 2.0 / (2 * wl["nmol"])
        EN = ddrax.spec_dens_extr_narr(tdev, G, "cartwright")     # Simpson on the device, 6 doubles to host
        return G, ddrax.dipolar(EN, "1H", "1H")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        e1.synchronize()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms.item()) / steps, out

    for _ in range(args.warmup):
        step(raw)
    sampler = ClockSampler(local) if rank == 0 else None
    lib.dynpy_profile_enable(1)
    lib.dynpy_profile_read(None, None)
    ms_step, (G, (rate, tc, F0)) = timed(lambda: step(raw), args.steps)
    import ctypes as C
    prof_ms = (C.c_double * 4)()
    prof_n = (C.c_int64 * 4)()
    lib.dynpy_profile_read(prof_ms, prof_n)
    lib.dynpy_profile_enable(0)
    clocks = sampler.stop() if sampler else None

    # ---- parity gate: no line is printed for a result that differs from the oracle
    check = None
    if ragged:
        counts = ops.dd_pair_count(wrapped, pi, pj, cell, rmax)
        present[0] = float(counts.sum().item())
        check = check_ragged(ops, wrapped, pi, pj, cell, rmax, F) if rank == 0 else None
    elif not args.no_check:
        check = check_result(ops, dist, wrapped, pi, pj, cell, F, M, G, 2 * wl["nmol"], lo, hi, rank)

    # ---- end to end through the public API with HOST buffers: what `python -m dynpy -d` does after the text parse
    #      (DDparse.dd_trajectory, reference DDparse.py:42-104).  Inside the timed region, every step: H2D of the raw
    #      coordinates of ALL atoms (O and H: the bond / molecule stage needs them) from pinned host memory, each rank
    #      moving 1/world over its own link; wrap; bonds + molecules of frame 0 and the device scan of EVERY frame
    #      against them (DD-3; frames sharded over the ranks); pair selection; the transforms; D2H of G(tau).
    del raw
    if not ragged:
        del wrapped
    torch.cuda.empty_cache()
    from dynpy_b200 import DDparse, topology
    A_all = box.nat
    cuts = [(A_all * r) // world for r in range(world + 1)]
    host_blk = torch.empty((cuts[rank + 1] - cuts[rank], 3, F), dtype=torch.float64, pin_memory=True)
    full = box.trajectory(F, chunk=2048)                       # all atoms, generated on the device
    host_blk.copy_(full[cuts[rank]:cuts[rank + 1]])
    del full
    torch.cuda.empty_cache()
    host_G = torch.empty((7, F), dtype=torch.float64, pin_memory=True)
    staging = torch.empty((A_all, 3, F), dtype=torch.float64, device=dev)
    wrapped_all = torch.empty_like(staging)
    topo_info = {}

    def e2e_step():
        mine = staging[cuts[rank]:cuts[rank + 1]]
        mine.copy_(host_blk, non_blocking=True)
        if world > 1:
            dist.gather_blocks(staging, cuts, mine)     # blocks exchanged GPU to GPU (NVLink)
        ops.wrap(staging, L, out=wrapped_all)
        # DD-3: the synthetic box has frames in which two waters touch inside the bond criterion; the scan lists them
        # (the product re-indexes those frames; with pair_type 'all' the pair list does not depend on them)
        ft, _ = topology.build_topology(wrapped_all, box.symbols, cell, rmax, bond_extra=0.45, dynamic=True,
                                        capacity=1 << 22)
        sp_i, sp_j = DDparse.select_pairs(box.symbols, ft.base, 3, wl["pair_type"], None)
        if args.pairs:
            sp_i, sp_j = sp_i[:args.pairs], sp_j[:args.pairs]
        assert len(sp_i) == P, (len(sp_i), P)
        topo_info.update(deviating_frames=len(ft.changes), molecules=ft.base.nmol)
        qi, qj = torch.from_numpy(sp_i).to(dev), torch.from_numpy(sp_j).to(dev)
        if ragged:
            Gd, hit, nspins, info = ddrax.pair_acf_sum(wrapped_all, qi, qj, cell, rmax)
            Gd = (Gd[:, hit] * (2.0 / nspins)).contiguous()
            r = ddrax.dipolar(ddrax.spec_dens_extr_narr(tdev[hit].contiguous(), Gd, "cartwright"), "1H", "1H")
        else:
            Gd = torch.zeros((7, F), dtype=torch.float64, device=dev)
            if hi > lo:
                spec = ops.dd_pairs_to_spectrum(wrapped_all, qi[lo:hi], qj[lo:hi], cell, M=M, ws_budget=ws_budget)
                Gd += ops.ifft_acf(spec, F, 1.0 / (float(M) * float(F)))
            dist.allreduce_sum_(Gd)
            Gd *= 2.0 / (2 * wl["nmol"])
            r = ddrax.dipolar(ddrax.spec_dens_extr_narr(tdev, Gd, "cartwright"), "1H", "1H")
        host_G[:, :Gd.shape[1]].copy_(Gd, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return Gd, r

    _, (rate_e2e, _, _) = e2e_step()
    if not args.pairs and abs(rate_e2e - rate) > 1e-9 * abs(rate):
        raise RuntimeError("bench: the end-to-end leg gives rate %r, the device-resident leg %r" % (rate_e2e, rate))
    ms_e2e, _ = timed(e2e_step, max(1, min(args.steps, 3)))

    if rank != 0:
        dist.shutdown()
        return
    hbm, how = peaks()
    pf_total = present[0]
    value = pf_total / (ms_step * 1e-3)
    names = ["dd_geom", "fft_cols", "fft_rows", "other"]
    per = {names[k]: {"ms_per_step": prof_ms[k] / args.steps, "launches_per_step": prof_n[k] / args.steps}
           for k in range(4)}
    dom = max(range(3), key=lambda k: prof_ms[k])
    pf_rank = float(hi - lo) * float(F) if not ragged else pf_total / world
    t_dom = prof_ms[dom] / args.steps * 1e-3
    achieved = ALG_BYTES_PER_PAIR_FRAME * pf_rank / t_dom / 1e9 if t_dom > 0 else 0.0
    # DRAM bytes per launch of the dominant kernel class from the committed ncu --set full capture
    # (profiles/traffic.json holds bytes per pair-frame; one launch = one batch of the workspace)
    traffic = None
    dram_per_pf = None
    launches = max(1.0, prof_n[dom] / args.steps)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj[names[dom]]["bytes_per_pair_frame"] * pf_rank / launches
        dram_per_pf = tj.get("whole_step_bytes_per_pair_frame")
    except Exception:
        pass
    alg_per_launch = ALG_BYTES_PER_PAIR_FRAME * pf_rank / launches
    line = {
        "metric": "pair-frames/s (dipolar G(tau)+R1)", "value": value, "unit": "pair-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "pairs": P, "frames": F, "fft_length": M, "cell_bohr": L,
                   "l2": "inputs larger than L2 (coords %.2f GB, per-batch FFT intermediates > 126 MB); no flush" %
                         (2 * wl["nmol"] * 3 * F * 8 / 1e9),
                   "sharding": "contiguous pair blocks per rank, one NCCL all-reduce of G[7,F]" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": pf_total / (ms_e2e * 1e-3), "unit": "pair-frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(A_all * 3 * F * 8), "d2h_bytes_per_step": int(7 * F * 8 + 48),
                "includes": "H2D of all atoms (sharded over the ranks), wrap, bonds + molecules of frame 0, device scan of "
                            "every frame against them (frames sharded over the ranks), pair selection, transforms, "
                            "all-reduce, Simpson + rate, D2H of G",
                "topology": topo_info},
        "gpu_launches": int(sum(prof_n)) + 2 * args.steps,
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "traffic": traffic, "algorithmic_bytes_per_launch": alg_per_launch,
                     "launch_ms": t_dom * 1e3 / launches, "peak_source": how,
                     "ceilings": {"note": "measured on this pool (profiles/r01_ubench_ceilings.txt): the path is bound by the fp64 pipe "
                                          "(row kernel 62 % busy) and the L1 / shared-memory data pipe (column kernel ~70 %), "
                                          "not by its 48 algorithmic bytes; the T intermediate (2 x 16 B per transform point) "
                                          "makes the DRAM traffic ~9 x the algorithmic bytes",
                                  "hbm_copy_gbs": 6200.0, "l2_copy_gbs": 12400.0, "fp64_lanes_per_clk_per_sm": 63.5,
                                  "dram_bytes_per_pair_frame": dram_per_pf},
                     "algorithmic_bytes_per_pair_frame": ALG_BYTES_PER_PAIR_FRAME,
                     "kernel_share_of_step": prof_ms[dom] / args.steps / ms_step,
                     "whole_step_frac": ALG_BYTES_PER_PAIR_FRAME * pf_rank / (ms_step * 1e-3) / 1e9 / hbm,
                     "per_kernel": per},
        "result": {"rate_1_over_T1_per_nH": rate, "tau_c_ps": tc, "G20_0": float(G[4, 0])},
        "check": check,
    }
    if ragged:
        line["config"]["rmax_bohr"] = rmax
        line["config"]["unit_note"] = ("pair-frames counted = pair-frames INSIDE rmax (%.4g of %.4g candidates): the "
                                       "reference only transforms those" % (pf_total, float(P) * float(F)))
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(box, wl, F)
    dist.shutdown()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ other mechanisms
def other_arm(args, wl):
    """Quadrupolar (cfg4 shape) and spin-rotation (cfg3) hot paths: same measurement contract as the dipolar arm
    (device-event timing, max over ranks, end-to-end leg with pinned host inputs, one JSON line)."""
    import torch
    import torch.distributed as tdist

    from dynpy_b200 import dist, ops, qrax, synth
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F = wl["frames"] if args.frames is None else args.frames
    if wl["kind"] == "qr":
        S = wl["nuclei"]  # per GPU: weak scaling, every rank owns its own nuclei (cfg4 shards 1e5 nuclei over 8 GPUs)
        if args.nuclei:
            S = args.nuclei
        base = synth.efg_series(250, F, seed=wl["seed"] + rank, device=dev)
        data = base.repeat((S + 249) // 250, 1, 1)[:S].contiguous()
        data += 1e-3 * torch.randn_like(data)
        units_rank, unit, alg_bytes, scaling = float(S) * F, "nucleus-frames/s", 48.0, "weak"
        tdev = torch.arange(F, device=dev, dtype=torch.float64) * 0.029
        M = ops.fft_length(F)

        nchunk = max(1, min(16, S // 256))
        cuts = [(S * k) // nchunk for k in range(nchunk + 1)]

        def step(src, ready=None):
            # `ready`: one event per chunk of nuclei (end-to-end leg: the chunk's host->device copy on the copy stream)
            spec = torch.zeros((3, M), dtype=torch.float64, device=dev)
            for k in range(nchunk):
                if ready is not None:
                    torch.cuda.current_stream().wait_event(ready[k])
                ops.qr_efg_to_spectrum(src[cuts[k]:cuts[k + 1]], M=M, spectrum=spec)
            dist.allreduce_sum_(spec)
            acf = ops.ifft_acf(spec, F, 1.0 / (float(M) * F * S * world))[[2, 1, 0, 1, 2]].contiguous()
            g, giso = qrax.spectral_dens(tdev, acf, "cartwright")
            return acf, qrax.relaxation(g, giso, "127I")
        metric = "nucleus-frames/s (quadrupolar ACFs + rates)"
    else:
        nmol = wl["nmol"]
        box = synth.methane_box(nmol, seed=wl["seed"]).to(dev)
        data = box.trajectory(F + 1, chunk=2048)
        mol_atoms = torch.arange(nmol * 5, dtype=torch.int32, device=dev).reshape(nmol, 5)
        mass = torch.tensor([12.010635561280681] + [1.0079407573975143] * 4, dtype=torch.float64, device=dev)
        axis_atoms = torch.stack([mol_atoms[:, 0], mol_atoms[:, 1], mol_atoms[:, 0], mol_atoms[:, 2]], 1).contiguous()
        L = box.celldm
        lo, hi = dist.shard_range(nmol, rank, world)
        units_rank, unit, alg_bytes, scaling = float(hi - lo) * F, "molecule-frames/s", 120.0, "strong"
        tdev = torch.arange(1, F + 1, device=dev, dtype=torch.float64) * 0.001
        M = ops.fft_length(F)

        nchunk = max(1, min(8, (hi - lo) // 16))
        cuts = [lo + ((hi - lo) * k) // nchunk // 2 * 2 for k in range(nchunk)] + [hi]   # even blocks: real-pair packing

        def step(src, ready=None):
            spec = torch.zeros((3, M), dtype=torch.float64, device=dev)
            for k in range(nchunk):
                if ready is not None:
                    torch.cuda.current_stream().wait_event(ready[k])
                a, b = cuts[k], cuts[k + 1]
                J, _, _ = ops.sr_angmom(src, mol_atoms[a:b].contiguous(), mass, axis_atoms[a:b].contiguous(), 0, (L, L, L),
                                        1000.0, want_omega=False, want_axes=False)
                for c in range(3):
                    ops.wk_acf_sum(J[:, c, :], pack_real_pairs=True, M=M, spectrum=spec[c])
            dist.allreduce_sum_(spec)
            acf = ops.ifft_acf(spec, F, 1.0 / (float(M) * F * nmol))
            return acf, ops.simpson(acf, tdev, "cartwright").cpu()
        metric = "molecule-frames/s (spin-rotation ACFs + rate)"

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        e1.synchronize()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms.item()) / steps, out

    for _ in range(args.warmup):
        step(data)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, _ = timed(lambda: step(data), args.steps)
    clocks = sampler.stop() if sampler else None
    # ---- end to end: the rank's inputs start in pinned host memory; they are copied chunk by chunk on a second
    #      stream while the chunks already on the device are transformed (the copy is the longer leg: PCIe)
    if wl["kind"] == "qr":
        rows = [(cuts[k], cuts[k + 1]) for k in range(nchunk)]          # nuclei of chunk k
    else:
        rows = [(5 * cuts[k], 5 * cuts[k + 1]) for k in range(nchunk)]  # atoms of the molecules of chunk k
    import psutil
    ring = data.numel() * 8 * world > 0.4 * psutil.virtual_memory().available   # all ranks pin host memory on one box
    if ring:   # not enough host memory to pin every rank's full input: one pinned chunk, re-sent for every chunk
        big = max(b - a for a, b in rows)
        host = torch.empty((big,) + tuple(data.shape[1:]), dtype=torch.float64, pin_memory=True)
        host.copy_(data[:big])
    else:
        host = torch.empty(data.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(data)
    staging = torch.empty_like(data)
    host_out = torch.empty((5 if wl["kind"] == "qr" else 3, F), dtype=torch.float64, pin_memory=True)
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_bytes = sum((b - a) for a, b in rows) * int(data[0].numel()) * 8

    def e2e_step():
        copy_stream.wait_stream(torch.cuda.current_stream())   # the previous step has consumed the staging buffer
        ready = []
        with torch.cuda.stream(copy_stream):
            for a, b in rows:
                staging[a:b].copy_(host[:b - a] if ring else host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append(ev)
        acf, r = step(staging, ready)
        host_out.copy_(acf, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return r

    e2e_step()
    ms_e2e, _ = timed(e2e_step, max(1, min(args.steps, 3)))
    if rank != 0:
        dist.shutdown()
        return
    hbm, how = peaks()
    total = units_rank * (world if scaling == "weak" else 1.0)
    if scaling == "strong":
        total = float(wl["nmol"]) * F
    line = {"metric": metric, "value": total / (ms_step * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "frames": F, "fft_length": M,
                       "l2": "inputs larger than L2 (%.2f GB per rank); no flush" % (data.numel() * 8 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(host_out.numel() * 8),
                    "overlap": "%d chunks copied on a second stream while earlier chunks are transformed" % nchunk,
                    "host_buffer": "one pinned chunk re-sent per chunk (host memory)" if ring else "full input pinned"},
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": alg_bytes * units_rank / (ms_step * 1e-3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": alg_bytes * units_rank / (ms_step * 1e-3) / 1e9 / hbm,
                         "traffic": None, "peak_source": how, "algorithmic_bytes_per_unit": alg_bytes}}
    dist.shutdown()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=None, help="override the frame count (debug)")
    ap.add_argument("--pairs", type=int, default=None, help="use only the first PAIRS pairs (debug)")
    ap.add_argument("--nuclei", type=int, default=None, help="quadrupolar workload: nuclei per GPU (default 12500)")
    ap.add_argument("--ws-gb", type=float, default=None,
                    help="workspace budget of the fused dipolar call in GiB (default: a quarter of the free memory, 4..32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle / Parseval gate (debug)")
    ap.add_argument("--rmax", type=float, default=None,
                    help="dipolar cut-off in bohr (default: the box length = gapless); 15 is the value of the "
                         "reference's notebooks/dynpy_params_water_dipolar.py and puts cfg2 on the ragged path")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if wl.get("kind") in ("qr", "sr"):
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                reference_arm_other(args, wl)
            return
        other_arm(args, wl)
    elif args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        reference_arm(args, wl)
    else:
        native_arm(args, wl)


if __name__ == "__main__":
    main()
