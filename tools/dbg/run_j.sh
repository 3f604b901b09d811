python -m pytest tests -m gpu -x -q -k "long_series or cfg5 or geometry_kernel_variants or fast_path" 2>&1 | tail -3
python tools/dbg/probe_dd.py 96 1000000 3 16 2>&1 | tail -1
python tools/dbg/probe_dd.py 192 500000 3 16 2>&1 | tail -1
python tools/dbg/probe_dd.py 384 250000 3 16 2>&1 | tail -1
