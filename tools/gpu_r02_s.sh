#!/bin/bash
# round-2 GPU call s (1 GPU): x backward with the merge in registers + transposed schedule -- parity, then A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft.py -x -q -m gpu > gpurun_out/r02_pairb_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pairb_parity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devptr" > gpurun_out/r02_pairb_parity2.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r02_pairb_parity2.log
ALT=$PWD/flutas_b200/csrc/libflutas_b200_nomerge.so
for rep in 1 2; do
  for w in NS C3 C5w1 C5xy; do
    for lib in pairb nomerge; do
      if [ $lib = nomerge ]; then export FLUTAS_B200_LIB=$ALT; else unset FLUTAS_B200_LIB; fi
      timeout 300 python bench.py --workload $w --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', d['config']['workload'][:4], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
    done
  done
done 2>&1 | tee gpurun_out/r02_pairb_ab.log
