#!/bin/bash
# round-2 GPU call h (1 GPU): y-kernel A/B builds, ncu launch list + full capture of the NS solver kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "matches_oracle and devptr and not generic" > gpurun_out/r02_pytest_h.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_h.log; tail -2 gpurun_out/r02_pytest_h.log
for v in "" _ldcs _ldcs_plainst; do
  for w in NS C2; do
    FLUTAS_B200_LIB=$PWD/flutas_b200/csrc/libflutas_b200$v.so python bench.py --workload $w --solver-only --steps 20 --warmup 5 --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib$v', '$w', d['value'], {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
  done
done | tee gpurun_out/r02_y_ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_NS.csv python bench.py --steps 2 --warmup 3 --no-parity > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"xfft_reg_kernel|yfft_reg_kernel|thomas_uni_tma|ref_solve|ref_scatter" -c 7 -f -o gpurun_out/r02_ns_full \
  python bench.py --solver-only --steps 1 --warmup 3 --no-parity > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_ns_full.ncu-rep gpurun_out/r02_launches_NS.csv
