#!/bin/bash
# round-2 GPU call b: full GPU test suite, source-level ncu capture of the NS y / z kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_b.log
grep -E "passed|failed|max\|dp\||rc=" gpurun_out/r02_pytest_gpu_b.log | tail -20
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"yfft_reg_kernel|thomas_uni_tma" -c 3 -f -o gpurun_out/r02_ns_src \
  python bench.py --solver-only --steps 1 --warmup 3 --no-parity > gpurun_out/r02_ncu_src.log 2>&1
ls -la gpurun_out/*.ncu-rep
