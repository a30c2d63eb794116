#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
run() { W=$1; tag=$2; shift 2; echo "== $W $tag" >> gpurun_out/ab.log; env "$@" timeout 300 python bench.py --solver-only --steps 10 --warmup 3 --workload $W 2>&1 | tail -1 >> gpurun_out/ab.log; }
CS=$PWD/flutas_b200/csrc
for W in ${WORKLOADS:-NS C2}; do
  run $W y8-cs FLUTAS_B200_Y8=1
  run $W y8-plain FLUTAS_B200_Y8=1 FLUTAS_B200_LIB=$CS/libflutas_b200_yplain.so
  run $W y16-plain FLUTAS_B200_Y8=0 FLUTAS_B200_LIB=$CS/libflutas_b200_yplain.so
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
