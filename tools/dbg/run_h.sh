for sub in 0 48 96 192 384; do echo "SUB=$sub"; DYNPY_TPC_SUB=$sub python tools/dbg/probe_dd.py 8192 100000 3 32 2>&1 | tail -1; done
