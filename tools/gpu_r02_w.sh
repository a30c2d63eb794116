#!/bin/bash
# round-2 GPU call w (1 GPU): y fwd on 16-lane tiles (new default) -- fft parity, default bench; A/B: x fwd with next-line loads issued under the pair pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r02_w_parity.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_w_parity.log
XFA=$PWD/flutas_b200/csrc/libflutas_b200_xfa.so
run() {  # label, lib, workload
  if [ -n "$2" ]; then export FLUTAS_B200_LIB=$2; else unset FLUTAS_B200_LIB; fi
  timeout 300 python bench.py --workload $3 --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'][:4], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
}
for rep in 1 2; do
  for w in NS C5w1 C5xy; do
    run "main " "" $w
    run "xfa  " "$XFA" $w
  done
done 2>&1 | tee gpurun_out/r02_w_ab.log
unset FLUTAS_B200_LIB
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_w.json 2> gpurun_out/r02_bench_NS_w.err; echo "bench rc=$?"
grep -a "^{" gpurun_out/r02_bench_NS_w.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_pressure_step'], d['parity']['err'], d['parity']['ok'], d['roofline']['frac'], d['roofline'].get('kernel'), d['roofline'].get('traffic'), {k:(v['ms'],v.get('frac')) for k,v in d['roofline']['stages'].items()}, d['e2e']['value'])"
