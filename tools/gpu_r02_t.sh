#!/bin/bash
# round-2 GPU call t (1 GPU): y kernels with the pair passes (16 values per thread) against the 8-value schedule and the old build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft.py -x -q -m gpu > gpurun_out/r02_ypair_parity.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_ypair_parity.log
FLUTAS_B200_Y8=0 timeout 600 python -m pytest tests/test_gpu_fft.py -x -q -m gpu > gpurun_out/r02_ypair_parity_y16.log 2>&1; echo "pytest y16 rc=$?"; tail -2 gpurun_out/r02_ypair_parity_y16.log
FLUTAS_B200_Y8=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devptr" > gpurun_out/r02_ypair_parity2.log 2>&1; echo "pytest2 rc=$?"; tail -2 gpurun_out/r02_ypair_parity2.log
YNOB=$PWD/flutas_b200/csrc/libflutas_b200_ynob.so
OLD=$PWD/flutas_b200/csrc/libflutas_b200_nomerge.so
run() {  # label, lib, y8, workload
  if [ -n "$2" ]; then export FLUTAS_B200_LIB=$2; else unset FLUTAS_B200_LIB; fi
  FLUTAS_B200_Y8=$3 timeout 300 python bench.py --workload $4 --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'][:4], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
}
for rep in 1 2; do
  for w in NS C5w1; do
    run "main-y8   " "" 1 $w
    run "main-y16pp" "" 0 $w
    run "ynob-y16p-" "$YNOB" 0 $w
    run "old-y16   " "$OLD" 0 $w
  done
  run "main  " "" 1 C5xy
  run "ynob  " "$YNOB" 1 C5xy
  run "old   " "$OLD" 1 C5xy
done 2>&1 | tee gpurun_out/r02_ypair_ab.log
