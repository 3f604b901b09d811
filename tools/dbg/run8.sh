#!/bin/bash
# 8-GPU lines of the three named multi-GPU configurations (cfg5 dipolar, cfg4 quadrupolar, cfg3 spin-rotation)
out=gpurun_out/r2; mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 8 --workload cfg5 --steps 2 --warmup 3 > $out/bench_cfg5_8gpu.json 2> $out/bench_cfg5_8gpu.err; echo "cfg5 rc=$?"; tail -c 600 $out/bench_cfg5_8gpu.err
timeout 300 $T bench.py --gpus 8 --workload cfg4 --steps 5 --warmup 3 > $out/bench_cfg4_8gpu.json 2> $out/bench_cfg4_8gpu.err; echo "cfg4 rc=$?"; tail -c 600 $out/bench_cfg4_8gpu.err
timeout 300 $T bench.py --gpus 8 --workload cfg3 --steps 5 --warmup 3 > $out/bench_cfg3_8gpu.json 2> $out/bench_cfg3_8gpu.err; echo "cfg3 rc=$?"; tail -c 600 $out/bench_cfg3_8gpu.err
cat $out/bench_cfg5_8gpu.json | cut -c1-400
