#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out; : > gpurun_out/pipe.log
FLUTAS_B200_PIPE=2 timeout 400 python -m pytest tests -m gpu -x -q -k "multi_gpu_slab" 2>&1 | tail -2
run() { W=$1; tag=$2; shift 2; echo "== $W $tag" >> gpurun_out/pipe.log; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --solver-only --steps 10 --warmup 3 2>&1 | grep '^{"metric"' | tail -1 >> gpurun_out/pipe.log; }
for W in ${WORKLOADS:-C2 NS}; do
  run $W off FLUTAS_B200_PIPE=0
  run $W c4x50 FLUTAS_B200_PIPE=4 FLUTAS_B200_PIPE_XSM=50
  run $W c4x35 FLUTAS_B200_PIPE=4 FLUTAS_B200_PIPE_XSM=35
  run $W c8x50 FLUTAS_B200_PIPE=8 FLUTAS_B200_PIPE_XSM=50
done
python - <<'PY'
import json
tag=None
for l in open('gpurun_out/pipe.log'):
    if l.startswith('=='): tag=l.strip(); continue
    try:
        d=json.loads(l); st=d['roofline']['stages']
        print("%-12s %7.2f Gpts/s %.3f ms "%(tag[3:], d['value'], d['ms_per_step']), " ".join("%s %.3f"%(k[:6]+k[-3:],v['ms']) for k,v in st.items()))
    except Exception as e: print(tag, 'ERR', l[:300])
PY
