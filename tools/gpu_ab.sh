#!/bin/bash
# A/B runs of kernel variants (environment variables / FLUTAS_B200_LIB builds); prints one summary line per run
mkdir -p gpurun_out
: > gpurun_out/ab.log
run() { W=$1; tag=$2; shift 2; echo "== $W $tag" >> gpurun_out/ab.log; env "$@" timeout 300 python bench.py --solver-only --steps 10 --warmup 3 --workload $W 2>&1 | tail -1 >> gpurun_out/ab.log; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "$(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "residual|FAILED|Error" gpurun_out/pytest_gpu.log | head

for W in ${WORKLOADS:-C5xy}; do
  run $W default X=1
  run $W ynarrow FLUTAS_B200_YWIDE=0
done
python - <<'PY'
import json
tag=None
for l in open('gpurun_out/ab.log'):
    if l.startswith('=='): tag=l.strip(); continue
    try:
        d=json.loads(l); st=d['roofline']['stages']
        print("%-16s %7.3f Gpts/s "%(tag[3:], d['value']), " ".join("%s %.3f (%.0f%%)"%(k[:6]+k[-3:],v['ms'],100*v.get('frac',0)) for k,v in st.items()))
    except Exception as e: print(tag, 'ERR', l[:300])
PY
