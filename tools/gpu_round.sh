#!/bin/bash
# One GPU call: parity tests, bench lines, ncu launch list and one full capture of every solver kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C2.log 2>&1; tail -1 gpurun_out/bench_C2.log
timeout 300 python bench.py --workload NS --solver-only --steps 5 --warmup 3 > gpurun_out/bench_NS.log 2>&1; tail -1 gpurun_out/bench_NS.log
timeout 300 python bench.py --workload C3 --solver-only --steps 10 --warmup 3 > gpurun_out/bench_C3.log 2>&1; tail -1 gpurun_out/bench_C3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fft|thomas' -s 15 -c 5 -o gpurun_out/prof_C2 -f python bench.py --steps 2 --warmup 3 --solver-only > gpurun_out/ncu_full_C2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fft|thomas' -s 15 -c 5 -o gpurun_out/prof_NS -f python bench.py --workload NS --steps 2 --warmup 3 --solver-only > gpurun_out/ncu_full_NS.log 2>&1
ls -la gpurun_out
