out=gpurun_out/r2; mkdir -p $out
python tools/dbg/probe_sr_shard.py 8 2>&1 | tail -3
python tools/dbg/probe_sr_shard.py 1 2>&1 | tail -3
for v in "" dynpy_b200/csrc/ab/alias2.so dynpy_b200/csrc/ab/alias1.so; do
  echo "LIB=$v"; DYNPY_B200_LIB=$v python tools/dbg/probe_dd.py 2000 100000 3 8 2>&1 | tail -1
done
