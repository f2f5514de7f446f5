#!/bin/bash
# full round job: parity tests, bench (both arms), ncu launch list + full captures of the three kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
nproc >> gpurun_out/r1_smi.txt; lscpu | grep "Model name" >> gpurun_out/r1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r1_pytest.log
tail -3 gpurun_out/r1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
cat gpurun_out/r1_bench.json gpurun_out/r1_bench_ref.json
timeout 300 python tools/prof_target.py --scheme selfsync --mib 100 --reps 3 > gpurun_out/plain_ss.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r1_launches_selfsync_100m.csv \
    python tools/prof_target.py --scheme selfsync --mib 100 --reps 3 > gpurun_out/ncu_ss.log 2>&1
timeout 300 python tools/prof_target.py --scheme selfsync --mib 100 --reps 2 > gpurun_out/plain_ss2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -s 2 -c 2 -o gpurun_out/r1_selfsync_100m \
    python tools/prof_target.py --scheme selfsync --mib 100 --reps 2 > gpurun_out/ncu_ss2.log 2>&1
ls -la gpurun_out | tail -8
