python tools/dbg/probe_topo.py 256 100000 2>&1 | tail -3
python -m pytest tests -m gpu -x -q -k "topolog or dynamic or cli or sr_ or neighbor or nn_" 2>&1 | tail -3
