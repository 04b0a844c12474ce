#!/bin/bash
# A/B timing of library builds on ONE box: tools/ab.sh "<time_paths args>" lib1.so lib2.so ...
# Each library is timed twice, interleaved with the others (boxes differ by a few percent, and so do the first and the
# second minute on a box); prints total ms and K1 ms of every repetition.
#   python -m osml_b200.build --out=tools/ab/x.so -DOSML_TC_LD128=1      builds a variant
cd "$(dirname "$0")/.."
args="$1"; shift
for round in 1 2; do
  for lib in "$@"; do
    echo "== $lib (round $round)"
    OSML_KNN_LIB=$PWD/$lib timeout 300 python tools/time_paths.py $args 2>&1 | grep -o "tensor n=.*ms \|last_filter_ms': [0-9.]*" | paste - - | tail -3
  done
done
