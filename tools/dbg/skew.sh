for sk in 0 1000 2000 3000; do echo "skew $sk"; DYNPY_ROWS_SKEW=$sk timeout 200 python tools/dbg/dbg_dd2.py 2000 100000 v2 | head -1; done
