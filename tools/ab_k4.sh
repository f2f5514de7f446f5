#!/bin/bash
# one gpurun job: fast parity subset, then device-timed decode throughput (both schemes, a few geometries and redundancies)
set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not GiB and not beyond" > gpurun_out/ab_tests.log 2>&1
tail -5 gpurun_out/ab_tests.log
python tools/quick_bench.py --mib 100 --r 0.5 --iters 10 --geom 0,0 2,768 2,640 1,864 > gpurun_out/ab_new.log 2>&1
python tools/quick_bench.py --mib 1024 --r 0.0 0.5 0.9 1.0 --iters 5 --schemes gap > gpurun_out/ab_new_1g.log 2>&1
cat gpurun_out/ab_new.log gpurun_out/ab_new_1g.log
