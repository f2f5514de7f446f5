#!/bin/bash
mkdir -p gpurun_out
python tools/prof_target.py --scheme gap --mib 100 --reps 2 --segs 2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_decode -s 1 -c 1 -o gpurun_out/r1_k4v2_segs2 \
    python tools/prof_target.py --scheme gap --mib 100 --reps 2 --segs 2 > gpurun_out/ncu3.log 2>&1
python tools/prof_target.py --scheme gap --mib 100 --reps 2 --segs 1 > gpurun_out/plain3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_decode -s 1 -c 1 -o gpurun_out/r1_k4v2_segs1 \
    python tools/prof_target.py --scheme gap --mib 100 --reps 2 --segs 1 > gpurun_out/ncu3b.log 2>&1
tail -2 gpurun_out/ncu3.log gpurun_out/ncu3b.log
