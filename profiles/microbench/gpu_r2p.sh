#!/bin/bash
# round 2, GPU call P: opacity write-out and source update without per-element index divisions
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2p_pytest.log
timeout 900 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2" --iters 3 --out gpurun_out/r2p_probe_16m.jsonl > gpurun_out/r2p_probe_16m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2p_probe_16m.jsonl
