#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/gputest_final.log
bash tools/dbg/run_final1.sh 2>&1 | tail -12
python bench.py --workload cfg5 --steps 1 --warmup 3 > $out/bench_cfg5_1gpu.json 2> $out/bench_cfg5_1gpu.err; tail -c 300 $out/bench_cfg5_1gpu.err; cut -c1-300 $out/bench_cfg5_1gpu.json
