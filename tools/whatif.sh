#!/bin/bash
# Developer timing experiments: K4 with parts of the work removed (results are WRONG on purpose; only the time matters).
# Variants are built by hand with -DHD_WHATIF=<mask> into build/libhd_w<mask>.so.
for w in "$@"; do
  echo "== HD_WHATIF=$w"
  HD_B200_LIB=$(pwd)/build/libhd_w$w.so python tools/quick_bench.py --mib 100 --r 0.5 --iters 5 --schemes ${SCHEME:-gap} 2>&1 | tail -1
done
