python -m pytest tests -m gpu -x -q -k "pair_count or ragged or dd_gaps or smoke" 2>&1 | tail -3
for gs in 0 1; do echo "M1=128 GS=$gs"; DYNPY_TPC3_GS=$gs python tools/dbg/probe_dd.py 384 250000 3 16 2>&1 | tail -1; done
python bench.py --rmax 15 --pairs 16384 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('ragged16k', d['value'], d['ms_per_step'], d['check'])"
