#!/bin/bash
# Round evidence in one gpurun job: GPU parity suite, both bench arms, the ncu launch list of bench.py itself, and full ncu
# captures of the self-sync decode (K1a + K4) and the gap-array decode (K4) at 1 GiB.  Everything lands in gpurun_out/.
R=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_smi.txt 2>&1
nproc >> gpurun_out/${R}_smi.txt; lscpu | grep "Model name" >> gpurun_out/${R}_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest.log
tail -3 gpurun_out/${R}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_ref.err; echo "reference arm rc=$?"
# launch list of the bench command itself (after it exited 0 without ncu above); short variant of the same command
timeout 600 python bench.py --steps 2 --warmup 1 --no-records --no-cpu-baseline > gpurun_out/plain_bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-records --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full captures (each after the same command exited 0 without ncu)
timeout 300 python tools/prof_target.py --scheme selfsync --mib 1024 --reps 2 > gpurun_out/plain_ss.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -s 2 -c 2 -o gpurun_out/${R}_selfsync_1g \
    python tools/prof_target.py --scheme selfsync --mib 1024 --reps 2 > gpurun_out/ncu_ss.log 2>&1
timeout 300 python tools/prof_target.py --scheme gap --mib 1024 --reps 2 > gpurun_out/plain_gap.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -s 1 -c 1 -o gpurun_out/${R}_gap_1g \
    python tools/prof_target.py --scheme gap --mib 1024 --reps 2 > gpurun_out/ncu_gap.log 2>&1
ls -la gpurun_out | tail -12
