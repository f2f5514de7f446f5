#!/bin/bash
# Developer timing experiments: K4 with parts of the work removed (results are WRONG on purpose; only the time matters).
# Variants are built by hand with -DHD_WHATIF=<mask> into build/libhd_w<mask>.so
# (K4: 1 = no look-back, 2 = no compaction, 4 = no decode loop, 8 = no flush, 64 = the decode loop stores its words
# lane-contiguously like a direct placement would; 66 = the estimate for a count-first K4;
# K1a: 256 = no parse, 512 = no in-warp repair; 2048 = jumps from the u16 table with explicit address arithmetic (the
# loop before the byte table), 4096 = every tile staged by the generic copy loop -- the last two are CORRECT builds, kept
# for A/B timing).  MIB / SCHEME select the workload.
for w in "$@"; do
  echo "== HD_WHATIF=$w"
  HD_B200_LIB=$(pwd)/build/libhd_w$w.so python tools/quick_bench.py --mib ${MIB:-1024} --r 0.5 --iters 5 --schemes ${SCHEME:-gap} 2>&1 | tail -2
done
