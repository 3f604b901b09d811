out=gpurun_out/r2; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/gputest5.log
python tools/dbg/probe_sr_shard.py 1 2>&1 | tail -2
python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/bench_cfg3_1gpu_v2.json 2> $out/bench_cfg3_1gpu_v2.err; tail -c 300 $out/bench_cfg3_1gpu_v2.err; cut -c1-330 $out/bench_cfg3_1gpu_v2.json
python bench.py --workload cfg4 --steps 5 --warmup 3 > $out/bench_cfg4_1gpu_v2.json 2> $out/bench_cfg4_1gpu_v2.err; tail -c 300 $out/bench_cfg4_1gpu_v2.err; cut -c1-330 $out/bench_cfg4_1gpu_v2.json
python tools/dbg/probe_topo.py 256 100000 2>&1 | tail -3
