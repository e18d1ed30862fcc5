#!/bin/bash
# round 2, GPU call N (8 GPUs): J through peer memory against the NCCL reduce-scatter at N = 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n_bench_16m_n8_peers.json 2> gpurun_out/r2n_bench_16m_n8_peers.err; echo "bench n8 peers rc=$?"; cat gpurun_out/r2n_bench_16m_n8_peers.json | cut -c1-3000; grep "peer-memory\|Error\|error" gpurun_out/r2n_bench_16m_n8_peers.err | head -n 5
VRT_NO_PEER_REDUCE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2n_bench_16m_n8_nccl.json 2> gpurun_out/r2n_bench_16m_n8_nccl.err; echo "bench n8 nccl rc=$?"; cat gpurun_out/r2n_bench_16m_n8_nccl.json | cut -c1-400
