#!/bin/bash
# round-2 GPU call y (1 GPU): why is the transposed backward schedule slower on periodic x lines?  ncu of x inv, both builds
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PM2=$PWD/flutas_b200/csrc/libflutas_b200_pm2.so
timeout 600 ncu --set full --import-source on --clock-control none -k regex:xfft_reg_kernel -s 6 -c 2 -f -o gpurun_out/r02_xinv_old \
  python bench.py --workload C3 --solver-only --steps 1 --warmup 3 --no-parity > gpurun_out/r02_ncu_xinv_old.log 2>&1
FLUTAS_B200_LIB=$PM2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:xfft_reg_kernel -s 6 -c 2 -f -o gpurun_out/r02_xinv_T \
  python bench.py --workload C3 --solver-only --steps 1 --warmup 3 --no-parity > gpurun_out/r02_ncu_xinv_T.log 2>&1
ls -la gpurun_out/r02_xinv_old.ncu-rep gpurun_out/r02_xinv_T.ncu-rep
