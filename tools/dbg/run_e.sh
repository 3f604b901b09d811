out=gpurun_out/r2; mkdir -p $out
N="ncu --set full --clock-control none --import-source on -f"
DYNPY_ROWS_GENERIC=0 $N -k regex:"k_rows" -s 40 -c 8 -o $out/ncu_rows_new python bench.py --rmax 15 --pairs 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-check > $out/ncu_rows_new.log 2>&1
DYNPY_ROWS_GENERIC=1 $N -k regex:"k_rows" -s 40 -c 8 -o $out/ncu_rows_old python bench.py --rmax 15 --pairs 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-check > $out/ncu_rows_old.log 2>&1
ls -la $out/ncu_rows_*.ncu-rep
