#!/bin/bash
# round 2, GPU call A: GPU tests in both visit orders, then the order probe at 16 M sites
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
VRT_BLOCKS=3,2 VRT_SLAB=5 VRT_STEP_MIN=64 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties_large.py -m gpu -x -q > gpurun_out/r2a_pytest_blocked.log 2>&1; echo "blocked rc=$?"
timeout 1500 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;1,1,0,2,VRT_EXPERIMENT=4;2,2,0,2;4,4,0,2;8,8,0,2;4,4,300,2;8,8,150,2;16,16,100,2;4,4,0,1" --out gpurun_out/r2a_order_16m.jsonl > gpurun_out/r2a_order_16m.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/r2a_pytest.log gpurun_out/r2a_pytest_blocked.log; cat gpurun_out/r2a_order_16m.jsonl
