#!/bin/bash
# one full ncu capture of the kernels of the second self-sync decode (K1a, K4); $1 = report name, $2 = MiB, $3 = redundancy
name=${1:-r2_selfsync_100m}
mib=${2:-100}
r=${3:-0.5}
mkdir -p gpurun_out
timeout 300 python tools/prof_target.py --scheme selfsync --mib $mib --r $r --reps 2 > gpurun_out/plain_ss2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -s 2 -c 2 -o gpurun_out/$name \
    python tools/prof_target.py --scheme selfsync --mib $mib --r $r --reps 2 > gpurun_out/ncu_ss2.log 2>&1
tail -2 gpurun_out/ncu_ss2.log
