#!/bin/bash
# round-2 GPU call u (1 GPU): full GPU suite, smoke and default bench on the tree with the pair-pass transforms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_u.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_u.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_u.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_u.json 2> gpurun_out/r02_bench_NS_u.err; echo "bench rc=$?"
grep -a "^{" gpurun_out/r02_bench_NS_u.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_pressure_step'], d['parity']['err'], d['parity']['ok'], d['roofline']['frac'], d['roofline'].get('kernel'), {k:(v['ms'],v.get('frac')) for k,v in d['roofline']['stages'].items()}, d['e2e']['value'], d['clocks'])"
for w in C3 C5w1 C2; do timeout 300 python bench.py --workload $w --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" >> gpurun_out/r02_solver_only_1gpu_u.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/r02_solver_only_1gpu_u.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:4], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})
PY
