#!/bin/bash
# round-2 GPU call q (1 GPU): main build (x buffer swizzled at N = 2048 only) vs swizzle everywhere with un-hoisted address XORs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft.py -x -q -m gpu > gpurun_out/r02_swz2_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_swz2_parity.log
ALT=$PWD/flutas_b200/csrc/libflutas_b200_swzo.so
FLUTAS_B200_LIB=$ALT timeout 600 python -m pytest tests/test_gpu_fft.py -x -q -m gpu > gpurun_out/r02_swz2_parity_alt.log 2>&1; echo "pytest alt rc=$?"; tail -3 gpurun_out/r02_swz2_parity_alt.log
for rep in 1 2; do
  for w in NS C3 C5w1 C5xy C2; do
    for lib in main swzo; do
      if [ $lib = swzo ]; then export FLUTAS_B200_LIB=$ALT; else unset FLUTAS_B200_LIB; fi
      timeout 300 python bench.py --workload $w --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', d['config']['workload'][:4], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
    done
  done
done 2>&1 | tee gpurun_out/r02_swz2_ab.log
