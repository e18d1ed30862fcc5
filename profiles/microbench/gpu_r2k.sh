#!/bin/bash
# round 2, GPU call K: J accumulation folded into the next batch's opacity kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regular_line.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2k_pytest.log
timeout 900 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;1,1,0,2,VRT_NO_FUSED_J=1" --iters 3 --out gpurun_out/r2k_fusedJ_16m.jsonl > gpurun_out/r2k_fusedJ_16m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2k_fusedJ_16m.jsonl
