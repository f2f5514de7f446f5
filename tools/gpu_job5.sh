#!/bin/bash
mkdir -p gpurun_out
python tools/prof_target.py --scheme gap --mib 100 --reps 2 --segs 2 --nt 512 > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_decode -s 1 -c 1 -o gpurun_out/r1_k4v3 \
    python tools/prof_target.py --scheme gap --mib 100 --reps 2 --segs 2 --nt 512 > gpurun_out/ncu5.log 2>&1
echo done
