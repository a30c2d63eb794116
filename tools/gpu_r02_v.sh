#!/bin/bash
# round-2 GPU call v (1 GPU): ncu launch list + full capture of the pair-pass kernels (NS); A/B: 16-lane y tiles, x inv periodic transposed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_NS_v.csv python bench.py --steps 2 --warmup 3 --no-parity > gpurun_out/r02_launches_bench_v.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xfft_reg_kernel|yfft_reg_kernel|thomas_uni_tma|ref_solve|ref_scatter" -c 7 -f -o gpurun_out/r02_ns_full_v \
  python bench.py --solver-only --steps 1 --warmup 3 --no-parity > gpurun_out/r02_ncu_full_v.log 2>&1
ls -la gpurun_out/r02_ns_full_v.ncu-rep gpurun_out/r02_launches_NS_v.csv
PM2=$PWD/flutas_b200/csrc/libflutas_b200_pm2.so
run() {  # label, lib, env, workload
  if [ -n "$2" ]; then export FLUTAS_B200_LIB=$2; else unset FLUTAS_B200_LIB; fi
  env $3 timeout 300 python bench.py --workload $4 --solver-only --no-parity --steps 20 --warmup 5 2>/dev/null | grep -a "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'][:4], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
}
for rep in 1 2; do
  run "main     " "" "X=1" NS
  run "ywide    " "" "FLUTAS_B200_YWIDE=1" NS
  run "xinv-T   " "$PM2" "X=1" NS
  run "main     " "" "X=1" C5w1
  run "ywide    " "" "FLUTAS_B200_YWIDE=1" C5w1
done 2>&1 | tee gpurun_out/r02_v_ab.log
