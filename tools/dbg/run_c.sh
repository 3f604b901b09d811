out=gpurun_out/r2; mkdir -p $out
python bench.py --rmax 15 --pairs 16384 --steps 1 --warmup 1 --no-cpu-baseline > $out/ragged16k.json 2> $out/ragged16k.err; tail -c 300 $out/ragged16k.err; cut -c1-300 $out/ragged16k.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/ragged16k_launches.csv python bench.py --rmax 15 --pairs 16384 --steps 1 --warmup 1 --no-cpu-baseline > $out/ragged16k_ncu.log 2>&1
tail -2 $out/ragged16k_ncu.log | cut -c1-200
