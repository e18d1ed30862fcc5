#!/bin/bash
# round 2, GPU call D (2 GPUs): multi-GPU numerics tests, then the default bench at N = 2 (in-library NCCL, LPT balance, cell-sliced e2e)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2d_pytest_multigpu.log 2>&1; echo "multigpu pytest rc=$?"; tail -n 15 gpurun_out/r2d_pytest_multigpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_16m_n2.json 2> gpurun_out/r2d_bench_16m_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/r2d_bench_16m_n2.json; grep "\[bench\]" gpurun_out/r2d_bench_16m_n2.err | tail -n 6; tail -n 5 gpurun_out/r2d_bench_16m_n2.err
grep -h "NVLS\|Connected\|nranks\|NCCL version" gpurun_out/nccl_n2_rank0.log 2>/dev/null | head -n 8
