#!/bin/bash
# usage: tools/scale_bench.sh N [workloads...]   -- solver-only strong-scaling lines on N GPUs of one box (p2p exchange)
N=$1; shift
WL=${@:-C2 NS}
for w in $WL; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --solver-only --steps 10 --warmup 3 2>&1 | grep '^{"metric"' | tail -1
done
if [ -n "$WITH_NCCL" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload C2 --solver-only --exchange nccl --steps 10 --warmup 3 2>&1 | grep '^{"metric"' | tail -1
fi
