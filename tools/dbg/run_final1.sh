#!/bin/bash
# final single-GPU measurement set of the round: smoke, bench line (+ reference arm), ncu launch list of the same command,
# ncu --set full of the three kernels of the cfg2 step, ragged bench line
out=gpurun_out/r2f; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > $out/bench_cfg2.json 2> $out/bench_cfg2.err; tail -c 300 $out/bench_cfg2.err; cut -c1-260 $out/bench_cfg2.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_cfg2_reference.json 2> $out/bench_cfg2_reference.err; cut -c1-200 $out/bench_cfg2_reference.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/b_plain.json 2> $out/b_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
python bench.py --pairs 8192 --steps 1 --warmup 1 --no-cpu-baseline > $out/b8k.json 2> $out/b8k.err && \
ncu --set full --clock-control none --import-source on -f -k regex:"k_cols_tpc|k_rows2|k_dd_geom4" -s 9 -c 6 -o $out/ncu_full_cfg2 python bench.py --pairs 8192 --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_full.log 2>&1
python bench.py --rmax 15 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_cfg2_rmax15.json 2> $out/bench_cfg2_rmax15.err; tail -c 200 $out/bench_cfg2_rmax15.err; cut -c1-200 $out/bench_cfg2_rmax15.json
ls -la $out
