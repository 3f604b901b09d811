#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $T bench.py --gpus 8 --steps 3 --warmup 3 > $out/bench_cfg2_8gpu_v2.json 2> $out/bench_cfg2_8gpu_v2.err; echo "cfg2 rc=$?"; tail -c 300 $out/bench_cfg2_8gpu_v2.err; cut -c1-260 $out/bench_cfg2_8gpu_v2.json
