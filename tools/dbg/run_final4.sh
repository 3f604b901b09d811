#!/bin/bash
# regression + ragged line after the in-place / transpose epilogue of the inverse transforms
out=gpurun_out/r2f; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/gputest_final4.log
python bench.py --rmax 15 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_cfg2_rmax15_c.json 2>/dev/null; python -c "import json; d=json.load(open('$out/bench_cfg2_rmax15_c.json')); print('ragged full', d['value'], d['ms_per_step'], d['e2e']['value'], d['check'], d['roofline']['per_kernel']['fft_rows'])"
