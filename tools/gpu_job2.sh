#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest2.log
tail -4 gpurun_out/r1_pytest2.log
python tools/quick_bench.py --mib 100 1024 --r 0.5 --iters 10 --segs 1 2 4 2>&1 | tee gpurun_out/r1_quick2.log
python tools/quick_bench.py --mib 256 --r 0.0 0.9 1.0 --iters 5 --segs 1 2 --schemes gap 2>&1 | tee -a gpurun_out/r1_quick2.log
