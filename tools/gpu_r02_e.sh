#!/bin/bash
# round-2 GPU call e (8 GPUs, charged 8x: keep it short): slab tests at world 4, bench NS at N=8 and N=4 (slab parity + autotune),
# C5 2048x2048x1024 at N=8 with the fused and the copy-engine backward exchange
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k multi_gpu > gpurun_out/r02_pytest_multigpu_N4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu_N4.log
grep -E "SLAB_OK|MISMATCH|passed|failed|rc=" gpurun_out/r02_pytest_multigpu_N4.log | tail -6
timeout 400 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N8.json 2> gpurun_out/r02_bench_NS_N8.err
tail -c 2600 gpurun_out/r02_bench_NS_N8.json
timeout 400 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N4.json 2> gpurun_out/r02_bench_NS_N4.err
tail -c 1500 gpurun_out/r02_bench_NS_N4.json
for z in 0 1; do
  FLUTAS_B200_ZCOPY=$z FLUTAS_B200_PIPE=0 timeout 300 $TR --nproc-per-node 8 --master-port 2956$z bench.py --gpus 8 --workload C5 --solver-only --steps 10 --warmup 3 >> gpurun_out/r02_C5_N8_zcopy.jsonl 2>> gpurun_out/r02_C5_N8_zcopy.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_C5_N8_zcopy.jsonl'):
    try:
        d=json.loads(l); print('C5 N8', d['value'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})
    except Exception as e: print('bad line', e)
PY
