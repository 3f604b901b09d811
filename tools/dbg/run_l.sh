for v in "" dynpy_b200/csrc/ab/dual128x2.so dynpy_b200/csrc/ab/dual256x1.so dynpy_b200/csrc/ab/nodual128x2.so ""; do
  echo "LIB=$v"; DYNPY_B200_LIB=$v python tools/dbg/probe_dd.py 2000 100000 3 8 2>&1 | tail -1
done
