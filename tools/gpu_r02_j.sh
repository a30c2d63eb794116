#!/bin/bash
# round-2 GPU call j (2 GPUs): distributed z on general grids (slab tests), single-GPU regression of the general z kernel,
# stretched-grid benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k multi_gpu > gpurun_out/r02_pytest_multigpu_j.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu_j.log
grep -E "str-|uni-chan|SLAB_OK|MISMATCH|passed|failed|rc=|rror" gpurun_out/r02_pytest_multigpu_j.log | cut -c1-180 | tail -24
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -s -k "(matches_oracle and devptr and not generic) or golden or (fullsize and (C2 or gr2)) or stencils" > gpurun_out/r02_pytest_regress_j.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_regress_j.log
grep -E "gr=|passed|failed|rc=" gpurun_out/r02_pytest_regress_j.log | cut -c1-260 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_j.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke_j.log
for g in 0 2; do
  python bench.py --workload NS --gr $g --solver-only --steps 10 --warmup 3 --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N1 NS gr=$g', d['value'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2959$g bench.py --gpus 2 --workload C3 --gr $g --solver-only --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N2 C3 gr=$g', d['value'], {k:v for k,v in d['slab_schedule'].items() if k!='note'}, {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done | tee gpurun_out/r02_stretched_j.log
