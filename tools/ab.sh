#!/bin/bash
# A/B timing of library builds: tools/ab.sh "<time_paths args>" lib1.so lib2.so ...   (each run prints total ms and K1 ms)
cd "$(dirname "$0")/.."
args="$1"; shift
for lib in "$@"; do
  for dbg in 0 4; do
    echo "== $lib debug=$dbg"
    OSML_KNN_LIB=$PWD/$lib python tools/time_paths.py $args debug=$dbg 2>&1 | grep -o "tensor n=.*ms \|last_filter_ms': [0-9.]*" | paste - - | tail -2
  done
done
