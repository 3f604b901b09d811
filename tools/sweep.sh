#!/bin/bash
# usage: tools/sweep.sh "<bench args>" lib1 lib2 ...   (libs under tools/variants)
args="$1"; shift
for lib in "$@"; do
  DYNPY_B200_LIB=$PWD/tools/variants/$lib python bench.py $args --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']
print('$lib', 'ms/step %.1f' % d['ms_per_step'], 'value %.3e' % d['value'], {k:(round(v['ms_per_step'],1), v['launches_per_step']) for k,v in pk.items()})
"
done
