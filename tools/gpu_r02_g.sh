#!/bin/bash
# round-2 GPU call g (8 GPUs, charged 8x): distributed z solve at N=8 and N=4 (full bench lines with slab parity), C3 and C5 solver-only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N8_dz.json 2> gpurun_out/r02_bench_NS_N8_dz.err
tail -c 2400 gpurun_out/r02_bench_NS_N8_dz.json
timeout 400 $TR --nproc-per-node 4 --master-port 29582 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N4_dz.json 2> gpurun_out/r02_bench_NS_N4_dz.err
tail -c 1200 gpurun_out/r02_bench_NS_N4_dz.json
for w in C3 C5; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29583 bench.py --gpus 8 --workload $w --solver-only --steps 10 --warmup 3 >> gpurun_out/r02_N8_dz_solver_only.jsonl 2>> gpurun_out/r02_N8_dz_solver_only.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_N8_dz_solver_only.jsonl'):
    try:
        d=json.loads(l); print(d['config']['workload'][:5],'N8', d['value'], d.get('slab_schedule'), {k:v['ms'] for k,v in d['roofline']['stages'].items()})
    except Exception as e: pass
PY
