out=gpurun_out/r2; mkdir -p $out
for g in 1 2 0; do
DYNPY_ROWS_GENERIC=$g python bench.py --rmax 15 --pairs 16384 --steps 2 --warmup 1 --no-cpu-baseline > $out/ragged16k_g$g.json 2> $out/ragged16k_g$g.err; tail -c 300 $out/ragged16k_g$g.err; echo "GENERIC=$g"; python - <<PY
import json; d=json.load(open("$out/ragged16k_g$g.json")); print(d["value"], d["ms_per_step"], d["roofline"]["per_kernel"]["fft_rows"], d["check"])
PY
done
python -m pytest tests -m gpu -x -q -k "ragged or wk or ifft or acf or cli or pipeline" 2>&1 | tail -3
