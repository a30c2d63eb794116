#!/bin/bash
# round-2 GPU call n (1 GPU): x lines of N = 1024 with 8 values per thread (FLUTAS_B200_X8) -- parity, then A/B on NS and C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
FLUTAS_B200_X8=1 timeout 600 python -m pytest tests/test_gpu_fft.py -x -q -m gpu -k arrplan > gpurun_out/r02_x8_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_x8_parity.log
for x8 in 0 1 0 1; do
  for w in NS C3; do
    FLUTAS_B200_X8=$x8 timeout 300 python bench.py --workload $w --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('X8=$x8', d['config']['workload'][:3], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
  done
done 2>&1 | tee gpurun_out/r02_x8_ab.log
