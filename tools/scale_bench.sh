N=$1
for w in C2 NS; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --solver-only --steps 10 --warmup 3 2>&1 | grep '^{"metric"' | tail -1
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload C2 --solver-only --exchange nccl --steps 10 --warmup 3 2>&1 | grep '^{"metric"' | tail -1
