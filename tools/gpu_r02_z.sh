#!/bin/bash
# round-2 GPU call z (8 GPUs): final tree -- default bench at 8 and 4 GPUs, C3 / C5 / stretched NS solver-only at 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
timeout 400 $TR --nproc-per-node $n --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N${n}_z.json 2> gpurun_out/r02_bench_NS_N${n}_z.err
grep -a "^{" gpurun_out/r02_bench_NS_N${n}_z.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NS N$n', d['value'], d['ms_per_step'], d['parity']['err'], d['parity']['ok'], {k:v for k,v in d['slab_schedule'].items() if k!='note'}, d['ms_per_pressure_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done
rm -f gpurun_out/r02_N8_z.jsonl
for args in "--workload C3" "--workload C5" "--workload NS --gr 2"; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29630 bench.py --gpus 8 $args --solver-only --steps 10 --warmup 3 2>> gpurun_out/r02_N8_z.err | grep -a "^{" >> gpurun_out/r02_N8_z.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_N8_z.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:4], d['config']['z_grid'][:8], 'N8', d['value'], d['ms_per_step'], {k:v for k,v in d['slab_schedule'].items() if k!='note'}, {k:v['ms'] for k,v in d['roofline']['stages'].items()})
PY
