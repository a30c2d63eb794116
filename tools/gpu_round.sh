#!/bin/bash
# One GPU call: parity tests, bench lines, ncu launch list and one full capture of every solver kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C2.log 2>&1; tail -1 gpurun_out/bench_C2.log | cut -c1-400
for W in NS C3 C5w1; do timeout 300 python bench.py --workload $W --solver-only --steps 10 --warmup 3 > gpurun_out/bench_$W.log 2>&1; tail -1 gpurun_out/bench_$W.log | cut -c1-120; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fft|thomas' -s 15 -c 5 -o gpurun_out/prof_C2 -f python bench.py --steps 2 --warmup 3 --solver-only > gpurun_out/ncu_full_C2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fft|thomas' -s 15 -c 5 -o gpurun_out/prof_NS -f python bench.py --workload NS --steps 2 --warmup 3 --solver-only > gpurun_out/ncu_full_NS.log 2>&1
ls -la gpurun_out | head -30
