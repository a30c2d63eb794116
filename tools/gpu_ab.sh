#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "$(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head -5
for v in 1 0; do FLUTAS_B200_CORREC_VEC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); st=d['roofline']['stages']; print('vec=$v', d['value'], d['ms_per_pressure_step'], {k:(v['ms'],v['frac']) for k,v in st.items() if k in ('fillps','correc')})"; done
