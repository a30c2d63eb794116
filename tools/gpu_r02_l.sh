#!/bin/bash
# round-2 GPU call l (8 GPUs): final tree -- default bench at N=8 (slab parity, autotune), stretched C3 and NS at N=8 (distributed z, general kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N8_final.json 2> gpurun_out/r02_bench_NS_N8_final.err
grep -a "^{" gpurun_out/r02_bench_NS_N8_final.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NS N8', d['value'], d['ms_per_step'], d['parity']['err'], d['parity']['ok'], {k:v for k,v in d['slab_schedule'].items() if k!='note'}, d['ms_per_pressure_step'])"
for w in C3 NS; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --workload $w --gr 2 --solver-only --steps 10 --warmup 3 >> gpurun_out/r02_N8_stretched.jsonl 2>> gpurun_out/r02_N8_stretched.err
done
grep -a "^{" gpurun_out/r02_N8_stretched.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][:4], d['config']['z_grid'], 'N8', d['value'], {k:v for k,v in d['slab_schedule'].items() if k!='note'}, {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
