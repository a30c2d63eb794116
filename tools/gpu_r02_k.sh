#!/bin/bash
# round-2 GPU call k (1 GPU): the driver's round-end sequence on the final tree + the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_k.log; tail -3 gpurun_out/r02_pytest_gpu_k.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_k.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_k.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_k.json 2> gpurun_out/r02_bench_NS_k.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_NS_k.json
/usr/bin/time -v python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_NS_k.json 2> gpurun_out/r02_bench_ref_NS_k.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/r02_bench_ref_NS_k.json; grep -E "Elapsed|Maximum resident" gpurun_out/r02_bench_ref_NS_k.err
