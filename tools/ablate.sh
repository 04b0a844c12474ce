#!/bin/bash
# K1 ablations at C4 (results meaningless for debug != 0; read last_filter_ms): python tools/time_paths.py n d q k path reps key=value
cd "$(dirname "$0")/.."
for dbg in 0 1 2 3 4; do
  echo "== debug=$dbg"
  python tools/time_paths.py 1000000 128 100000 10 tensor 3 debug=$dbg 2>&1 | grep -o "tensor n=.*ms \|last_filter_ms': [0-9.]*" | paste - - | tail -2
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu && ./tools/tmem_bw
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu && ./tools/mma_rate
