#!/bin/bash
# one full ncu capture of the three kernels of the second decode (K1a, K1b, K4)
mkdir -p gpurun_out
timeout 300 python tools/prof_target.py --scheme selfsync --mib 100 --reps 2 > gpurun_out/plain_ss2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -s 2 -c 2 -o gpurun_out/r1_selfsync_100m \
    python tools/prof_target.py --scheme selfsync --mib 100 --reps 2 > gpurun_out/ncu_ss2.log 2>&1
tail -2 gpurun_out/ncu_ss2.log
