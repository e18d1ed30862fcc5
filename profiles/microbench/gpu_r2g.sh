#!/bin/bash
# round 2, GPU call G (8 GPUs): the default bench at N = 8 and N = 4 (in-library NCCL, LPT balance, deferred all-gather, cell-sliced e2e)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench_16m_n8.json 2> gpurun_out/r2g_bench_16m_n8.err; echo "bench n8 rc=$?"; cat gpurun_out/r2g_bench_16m_n8.json; grep "\[bench\]" gpurun_out/r2g_bench_16m_n8.err | tail -n 4; tail -n 3 gpurun_out/r2g_bench_16m_n8.err
grep -h "NVLS\|nranks\|NCCL version\|Channel\|nChannels" gpurun_out/nccl_n8_rank0.log 2>/dev/null | head -n 12
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2g_bench_16m_n4.json 2> gpurun_out/r2g_bench_16m_n4.err; echo "bench n4 rc=$?"; cat gpurun_out/r2g_bench_16m_n4.json; tail -n 3 gpurun_out/r2g_bench_16m_n4.err
