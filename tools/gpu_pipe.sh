#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out; : > gpurun_out/zcopy.log
FLUTAS_B200_ZCOPY=1 timeout 400 python -m pytest tests -m gpu -x -q -k "multi_gpu_slab" 2>&1 | tail -1
run() { W=$1; tag=$2; shift 2; echo "== $W $tag" >> gpurun_out/zcopy.log; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --solver-only --steps 5 --warmup 3 2>&1 | grep '^{"metric"' | tail -1 >> gpurun_out/zcopy.log; }
for W in ${WORKLOADS:-C5 NS}; do
  run $W fused FLUTAS_B200_ZCOPY=0
  run $W copy FLUTAS_B200_ZCOPY=1
done
python - <<'PY'
import json
tag=None
for l in open('gpurun_out/zcopy.log'):
    if l.startswith('=='): tag=l.strip(); continue
    try:
        d=json.loads(l); st=d['roofline']['stages']
        print("%-12s %7.2f Gpts/s %.3f ms "%(tag[3:], d['value'], d['ms_per_step']), " ".join("%s %.3f"%(k[:6]+k[-3:],v['ms']) for k,v in st.items()))
    except Exception as e: print(tag, 'ERR', l[:300])
PY
