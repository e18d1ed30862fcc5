#!/bin/bash
# round 2, GPU call E: the decoupled sweep kernel (ring without dependencies, intensities gathered by the consumers)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
VRT_SWEEP_DECOUPLED=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -x -q > gpurun_out/r2e_pytest_decoupled.log 2>&1; echo "decoupled pytest rc=$?"; tail -n 3 gpurun_out/r2e_pytest_decoupled.log
VRT_SWEEP_DECOUPLED=1 VRT_BLOCKS=3,2 VRT_SLAB=5 VRT_STEP_MIN=64 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2e_pytest_decoupled_blocked.log 2>&1; echo "decoupled+blocked pytest rc=$?"; tail -n 3 gpurun_out/r2e_pytest_decoupled_blocked.log
timeout 1500 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;1,1,0,2,VRT_SWEEP_DECOUPLED=1;1,1,0,2,VRT_SWEEP_DECOUPLED=1,VRT_TMA_CFG=18;1,1,0,2,VRT_SWEEP_DECOUPLED=1,VRT_TMA_SMEM_KB=75;2,2,0,2,VRT_SWEEP_DECOUPLED=1;4,4,0,2,VRT_SWEEP_DECOUPLED=1;8,8,0,2,VRT_SWEEP_DECOUPLED=1;4,4,300,2,VRT_SWEEP_DECOUPLED=1" --out gpurun_out/r2e_decoupled_16m.jsonl > gpurun_out/r2e_decoupled_16m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2e_decoupled_16m.jsonl
