#!/bin/bash
# round-2 GPU call f (2 GPUs): distributed z solve -- slab tests (all three modes), bench NS at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k multi_gpu > gpurun_out/r02_pytest_multigpu_dz_N2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multigpu_dz_N2.log
grep -E "slab p2p|slab nccl|SLAB_OK|MISMATCH|passed|failed|rc=|Error|error" gpurun_out/r02_pytest_multigpu_dz_N2.log | cut -c1-200 | tail -45
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_NS_N2_dz.json 2> gpurun_out/r02_bench_NS_N2_dz.err
tail -c 2800 gpurun_out/r02_bench_NS_N2_dz.json; tail -3 gpurun_out/r02_bench_NS_N2_dz.err | cut -c1-300
