#!/bin/bash
# last regression of the round: full GPU suite, smoke, default bench line
out=gpurun_out/r2f; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > $out/bench_cfg2_v6.json 2> $out/bench_cfg2_v6.err; tail -c 300 $out/bench_cfg2_v6.err; cut -c1-230 $out/bench_cfg2_v6.json
