#!/usr/bin/env python3
"""Instruction mix of the hot kernels from the shipped cubins (cuobjdump -sass), per kernel: total SASS
instructions, fp64 (DFMA / DADD / DMUL / DSETP), shared / global / local memory, bulk-async (TMA) and barrier
opcodes.   python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dynpy_b200", "csrc", "libdynpy_b200.so")
pat = sys.argv[1] if len(sys.argv) > 1 else r"k_rows2|k_cols_tpc|k_dd_geom|k_topology_check|k_sr_angmom|k_cols_tpcg.*QrLoader|k_dd_ragged_words|k_pdist"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, mix = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        mix[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        mix[name][m.group(1)] += 1
GROUPS = [("DFMA", r"^DFMA"), ("DADD", r"^DADD"), ("DMUL", r"^DMUL"), ("DSETP", r"^DSETP"), ("LDS", r"^LDS"), ("STS", r"^STS"),
          ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDL", r"^LDL"), ("STL", r"^STL"), ("LDGSTS", r"^LDGSTS"),
          ("UBLKCP/UTMA", r"^(UBLKCP|UTMA)"), ("UBLKPF", r"^UBLKPF"), ("SYNCS(mbarrier)", r"^SYNCS"), ("BAR", r"^BAR"),
          ("RED/ATOM", r"^(RED|ATOM)"), ("MOV-like", r"^(MOV|IMAD\.MOV|UMOV)"), ("SHFL", r"^SHFL")]
for k, c in mix.items():
    if not re.search(pat, k):
        continue
    tot = sum(c.values())
    short = re.sub(r"\(.*", "", k)
    print("%-70s total %6d  " % (short[:70], tot) + "  ".join("%s %d" % (g, sum(v for op, v in c.items() if re.match(rx, op)))
                                                              for g, rx in GROUPS if sum(v for op, v in c.items() if re.match(rx, op))))
