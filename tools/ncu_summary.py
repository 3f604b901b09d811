#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): one block per profiled launch with the metrics that
matter for this code base (duration, DRAM/L2/L1/LSU/fp64 utilisation, occupancy, stall mix)."""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_inst%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible/clk"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("sm__cycles_elapsed.avg", "cycles"), ("smsp__inst_executed.sum", "warp_inst"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("=====", d["Kernel Name"][:70])
        print("   ", "  ".join("%s=%s%s" % (n, d[k], ("" if u[k] in ("", "%") else " " + u[k])) for k, n in KEYS if k in d))
        st = []
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(v)))
                except ValueError:
                    pass
        st.sort(key=lambda x: -x[1])
        print("    stalls/issue:", "  ".join("%s=%.2f" % kv for kv in st[:8]))


if __name__ == "__main__":
    main()
