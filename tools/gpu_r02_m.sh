#!/bin/bash
# round-2 GPU call m (1 GPU): fft(plan,arr) / guru plans / FFTW seam library against the oracle; reference arm at NS on the box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fft.py tests/test_abi.py -x -q -m "gpu or not gpu" > gpurun_out/r02_pytest_gpu_fft.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu_fft.log
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$? wall=$(( $(date +%s) - t0 )) s"
tail -c 1500 gpurun_out/r02_bench_reference_arm.json
