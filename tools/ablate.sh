#!/bin/bash
# K1 ablations with cycle counters (results meaningless for debug != 0); SHAPE="n d q k" (default: C4)
cd "$(dirname "$0")/.."
for dbg in ${1:-0 1 2 3 4}; do
  echo "== debug=$dbg"
  OSML_KNN_LIB=$PWD/tools/ab/counters.so python tools/time_paths.py ${SHAPE:-1000000 128 100000 10} tensor 2 debug=$dbg counters=1 2>&1 | grep -o "tensor n=.*ms \|last_filter_ms': [0-9.]*\|\[K1 dbg\].*" | tail -5
done
