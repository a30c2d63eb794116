#!/bin/bash
# round-2 GPU call d (2 GPUs): bounduvw/chkdt kernels, multi-GPU slab tests (kept log), bench NS at N=2 with slab parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bounduvw.py -m gpu -q -x > gpurun_out/r02_pytest_bounduvw.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_bounduvw.log; tail -3 gpurun_out/r02_pytest_bounduvw.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k multi_gpu > gpurun_out/r02_pytest_multigpu_N2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu_N2.log
grep -E "slab|SLAB|passed|failed|rc=" gpurun_out/r02_pytest_multigpu_N2.log | tail -40
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N2.json 2> gpurun_out/r02_bench_NS_N2.err
tail -c 2500 gpurun_out/r02_bench_NS_N2.json; tail -3 gpurun_out/r02_bench_NS_N2.err
