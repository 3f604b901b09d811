# usage: sweep_env.sh VAR v1 v2 ... : run the dipolar probe (default kernels) once per value of VAR
var=$1; shift
for v in "$@"; do echo "$var=$v"; env $var=$v timeout 200 python tools/dbg/dbg_dd3.py 2000 100000 tpc 2>/dev/null | head -1; done
