#!/bin/bash
# round-1 GPU job: parity tests, bench (both arms), ncu launch list + full capture of the decode kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
nproc >> gpurun_out/r1_smi.txt; lscpu | grep "Model name" >> gpurun_out/r1_smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
python tools/quick_bench.py --mib 100 1024 --r 0.5 --iters 10 > gpurun_out/r1_quick.log 2>&1
python tools/prof_target.py --scheme selfsync --mib 100 --reps 3 > gpurun_out/plain_ss.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches_selfsync_100m.csv \
    python tools/prof_target.py --scheme selfsync --mib 100 --reps 3 > gpurun_out/ncu_ss.log 2>&1
python tools/prof_target.py --scheme selfsync --mib 100 --reps 2 > gpurun_out/plain_ss2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -s 4 -c 4 -o gpurun_out/r1_selfsync_100m \
    python tools/prof_target.py --scheme selfsync --mib 100 --reps 2 > gpurun_out/ncu_ss2.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_bench.json; cat gpurun_out/r1_bench_ref.json; cat gpurun_out/r1_quick.log
