#!/bin/bash
# round-2 GPU call c (1 GPU): new transform kinds + z-kernel change, NS / C3 / C5w1 solver timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or oracle or reference_order or unsupported or stencils" > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_c.log
tail -4 gpurun_out/r02_pytest_gpu_c.log
for w in NS C3 C5w1 C2; do python bench.py --workload $w --solver-only --steps 20 --warmup 5 --no-parity; done > gpurun_out/r02_solver_only_c.jsonl 2> gpurun_out/r02_solver_only_c.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_solver_only_c.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:4], d['value'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})
PY
