#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_target.py --scheme gap --mib 16 --reps 2 > gpurun_out/smoke6.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke6.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r1_pytest6.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r1_pytest6.log
timeout 200 python tools/trace_k4.py --mib 100 --segs 2 --nt 512 2>&1| head -12
timeout 300 python tools/quick_bench.py --mib 100 1024 --r 0.5 --iters 10 --geom 2,512 2>&1 | tee gpurun_out/r1_quick6.log
