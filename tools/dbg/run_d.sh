out=gpurun_out/r2; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/gputest6.log
for g in 1 0; do
DYNPY_ROWS_GENERIC=$g python bench.py --rmax 15 --pairs 16384 --steps 2 --warmup 1 --no-cpu-baseline > $out/ragged16k_g$g.json 2> $out/ragged16k_g$g.err; tail -c 300 $out/ragged16k_g$g.err; echo "GENERIC=$g"; python - <<PY
import json; d=json.load(open("$out/ragged16k_g$g.json")); print(d["value"], d["ms_per_step"], d["roofline"]["per_kernel"], d["check"])
PY
done
python bench.py --rmax 15 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_cfg2_rmax15_v2.json 2> $out/bench_cfg2_rmax15_v2.err; tail -c 300 $out/bench_cfg2_rmax15_v2.err; cut -c1-200 $out/bench_cfg2_rmax15_v2.json
