#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg2_v7.json 2> $out/bench_cfg2_v7.err; tail -c 200 $out/bench_cfg2_v7.err; python -c "import json; d=json.load(open('$out/bench_cfg2_v7.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['check'])"
