#!/bin/bash
# usage: ab_build.sh NAME "EXTRA nvcc flags"  -> dynpy_b200/csrc/ab/NAME.so (a variant build of the library for same-box
# A/B runs: DYNPY_B200_LIB=dynpy_b200/csrc/ab/NAME.so python tools/dbg/probe_dd.py ...)
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/../.." && pwd)
tmp=/tmp/ab_$name
rm -rf "$tmp"; mkdir -p "$tmp/dynpy_b200" "$tmp/include"
cp -r "$root/dynpy_b200/csrc" "$tmp/dynpy_b200/csrc"; cp -r "$root/include/." "$tmp/include/"
rm -rf "$tmp/dynpy_b200/csrc/build" "$tmp/dynpy_b200/csrc/libdynpy_b200.so" "$tmp/dynpy_b200/csrc/ab"
make -C "$tmp/dynpy_b200/csrc" -j8 EXTRA="$extra" > "$tmp/make.log" 2>&1 || { tail -20 "$tmp/make.log"; exit 1; }
mkdir -p "$root/dynpy_b200/csrc/ab"
cp "$tmp/dynpy_b200/csrc/libdynpy_b200.so" "$root/dynpy_b200/csrc/ab/$name.so"
grep -h -A2 "k_cols_tpcILi50\|k_rows2ILb0" "$tmp/dynpy_b200/csrc/build/dd2.ptxas.log" | grep "spill\|registers" | tr '\n' ' '; echo
