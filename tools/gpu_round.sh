#!/bin/bash
# One GPU call: parity tests, smoke, bench lines, ncu launch list and one full capture of every solver kernel
# (reports are reduced to CSV on the box: gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest: $(tail -1 gpurun_out/pytest_gpu.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
fi
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C2.log 2>&1; tail -1 gpurun_out/bench_C2.log | cut -c1-200
for W in NS C3 C5w1 C5xy; do timeout 300 python bench.py --workload $W --solver-only --steps 10 --warmup 3 > gpurun_out/bench_$W.log 2>&1; tail -1 gpurun_out/bench_$W.log | cut -c1-120; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'fft|thomas|correc|fillps' -s 15 -c 7 -o /tmp/prof_C2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_C2.log 2>&1
ncu -i /tmp/prof_C2.ncu-rep --page raw --csv > gpurun_out/prof_C2_raw.csv
timeout 400 ncu --set full --clock-control none -k regex:'fft|thomas' -s 15 -c 5 -o /tmp/prof_NS -f python bench.py --workload NS --steps 2 --warmup 3 --solver-only > gpurun_out/ncu_full_NS.log 2>&1
ncu -i /tmp/prof_NS.ncu-rep --page raw --csv > gpurun_out/prof_NS_raw.csv
du -sh gpurun_out
