#!/bin/bash
# round-2 GPU call x (2 GPUs): slab tests (NCCL all-to-all, direct stores + distributed z, direct stores + transposes) and the
# default bench on two GPUs with the pair-pass transforms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k multi_gpu > gpurun_out/r02_pytest_multigpu_x.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu_x.log
grep -E "SLAB_OK|MISMATCH|passed|failed|rc=|rror" gpurun_out/r02_pytest_multigpu_x.log | cut -c1-200 | tail -30
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N2_x.json 2> gpurun_out/r02_bench_NS_N2_x.err
grep -a "^{" gpurun_out/r02_bench_NS_N2_x.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NS N2', d['value'], d['ms_per_step'], d['parity']['err'], d['parity']['ok'], {k:v for k,v in d['slab_schedule'].items() if k!='note'}, {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
for w in C3 C5w1; do
timeout 300 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --workload $w --solver-only --steps 10 --warmup 3 2>/dev/null | grep -a "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N2', d['config']['workload'][:4], d['value'], d['parity']['err'] if 'parity' in d and d['parity'] else None, {k:v for k,v in d['slab_schedule'].items() if k!='note'}, {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done
