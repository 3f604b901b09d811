#!/bin/bash
# ncu --set full captures of the kernels outside the cfg2 dipolar step (one GPU; each after the plain command ran):
# spin-rotation (k_sr_angmom), quadrupolar column pass (k_cols_tpcg<32, QrLoader>), the long-column kernel of cfg5
# (k_cols_tpc3), the ragged path (presence count, word kernel, row kernel with write epilogues, transposes, scatter)
# and the topology scan.
set -x
out=gpurun_out/r2f; mkdir -p $out
N="ncu --set full --clock-control none --import-source on -f"
python tools/bench_other.py 1024 > $out/other_plain.log 2>&1 && tail -4 $out/other_plain.log && \
$N -k regex:"k_sr_angmom|k_cols_tpcg" -c 3 -o $out/ncu_other python tools/bench_other.py 1024 > $out/ncu_other.log 2>&1
python tools/dbg/probe_dd.py 96 1000000 1 16 > $out/cfg5_plain.log 2>&1 && \
$N -k regex:"k_cols_tpc3" -s 1 -c 1 -o $out/ncu_cfg5 python tools/dbg/probe_dd.py 96 1000000 1 16 > $out/ncu_cfg5.log 2>&1
python tools/dbg/dbg_ragged.py 2000 > $out/ragged_plain.log 2>&1 && tail -3 $out/ragged_plain.log && \
$N -k regex:"k_dd_ragged_words|k_spectrum_even|k_dd_ragged_scatter|k_dd_pair_count|k_rows2" -s 2 -c 9 -o $out/ncu_ragged python tools/dbg/dbg_ragged.py 2000 > $out/ncu_ragged.log 2>&1
python tools/dbg/probe_topo.py 256 20000 > $out/topo_plain.log 2>&1 && \
$N -k regex:"k_topology_check" -c 1 -o $out/ncu_topo python tools/dbg/probe_topo.py 256 20000 > $out/ncu_topo.log 2>&1
ls -la $out/*.ncu-rep
