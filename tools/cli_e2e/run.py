#!/usr/bin/env python
"""CLI-level end-to-end timing of `python -m dynpy -d` on a cfg2-shaped .xyz (SURVEY §8d asks for the hot path;
this tool times what a user of the drop-in actually runs: text parse, H2D, bonds/molecules of every frame, pair
selection, transforms, CSV).   python tools/cli_e2e/run.py [nmol] [frames] [rmax]  -> one JSON line."""
import json, os, subprocess, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dynpy_b200 import synth
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rmax = float(sys.argv[3]) if len(sys.argv) > 3 else None
work = tempfile.mkdtemp(prefix="dynpy_cli_e2e_")
os.makedirs(os.path.join(work, "01"))
# a calm box (small translations): the default synthetic box of the bench lets neighbouring molecules touch inside
# the reference's bond criterion a few times per frame, which the driver reports as "too reactive"
box = synth.water_box(nmol, seed=2, trans_amp=float(os.environ.get("CLI_E2E_TRANS_AMP", "0.03"))).to("cuda")
traj = box.trajectory(F, chunk=4096)                       # [A,3,F] bohr
ang = (traj / synth.BOHR_PER_ANG).permute(2, 0, 1).contiguous().cpu().numpy()
ang.tofile(os.path.join(work, "traj.bin"))
with open(os.path.join(work, "sym.txt"), "w") as fh:
    fh.write("\n".join(box.symbols) + "\n")
exe = os.path.join(work, "xyz_writer")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tools", "cli_e2e", "xyz_writer.cpp")])
t0 = time.time()
subprocess.check_call([exe, os.path.join(work, "traj.bin"), str(box.nat), str(F), os.path.join(work, "sym.txt"),
                       os.path.join(work, "01", "traj.xyz")])
t_write = time.time() - t0
os.remove(os.path.join(work, "traj.bin"))
L = box.celldm
rm = L if rmax is None else rmax
synth.write_params(os.path.join(work, "params.py"), """class ParseDynamics:
    MD_format = 'xyz'
    traj_dir = './'
    trajs = ['01']
    nat = %d
    start_prod = 1
    end_prod = %d
    md_print_freq = 1
    sample_freq = 1
    celldm = %r
    timestep = 0.001
class Dipolar:
    identifier = 'H(2)O(1)'
    pair_type = 'all'
    rmax = %r
""" % (box.nat, F, L, rm))
del traj, ang
torch.cuda.empty_cache()
env = dict(os.environ)
env["PYTHONPATH"] = ROOT
env["DYNPY_STAGE_TIMES"] = os.path.join(work, "stages.json")
size = os.path.getsize(os.path.join(work, "01", "traj.xyz"))
runs = []
for rep in range(2):      # the second run reads the file from the page cache and finds the CUDA context warm-up paid
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "dynpy", "-d", "params.py"], cwd=work, input="y\n", text=True,
                       capture_output=True, env=env)
    wall = time.time() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:], file=sys.stderr)
        sys.exit(1)
    st = json.load(open(os.path.join(work, "stages.json")))
    st["process_wall_s"] = wall
    st["result_line"] = [l for l in r.stdout.split("\n") if "1/T1/nH" in l][:1]
    runs.append(st)
best = runs[-1]
pf = best["pairs"] * best["frames"]
out = {"tool": "cli_e2e", "workload": "python -m dynpy -d, %d H2O (%d atoms) x %d frames .xyz, pair_type all, rmax %.4g" % (nmol, box.nat, F, rm),
       "xyz_bytes": size, "xyz_write_s_tool": t_write, "runs": runs,
       "driver_pair_frames_per_s": pf / best["total_s"], "process_pair_frames_per_s": pf / best["process_wall_s"],
       "parse_GBps_text": size / 1e9 / best["stages_s"]["trajectory text -> device [A][3][F] (parse + H2D + transpose)"]}
print(json.dumps(out))
import shutil
shutil.rmtree(work, ignore_errors=True)
