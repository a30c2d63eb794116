#!/bin/bash
# round-2 GPU call i (1 GPU): y 16-lane RR=8 variant A/B; rehearsal of the driver's round-end sequence (pytest -m gpu, smoke, bench, bench --impl reference at C2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 0 1; do
  FLUTAS_B200_Y8WIDE=$v python bench.py --workload NS --solver-only --steps 20 --warmup 5 --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('y8wide=$v NS', d['value'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done | tee gpurun_out/r02_y8wide_ab.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_i.log; tail -3 gpurun_out/r02_pytest_gpu_i.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_i.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_i.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_i.json 2> gpurun_out/r02_bench_NS_i.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_NS_i.json
