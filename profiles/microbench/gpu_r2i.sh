#!/bin/bash
# round 2, GPU call I: bounded L2 prefetch of the visit rows ahead of the stage ring
cd "${GRAFT_REPO_ROOT:-/root/repo}"
VRT_PF_LEAD=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2i_pytest_prefetch.log 2>&1; echo "prefetch pytest rc=$?"; tail -n 3 gpurun_out/r2i_pytest_prefetch.log
timeout 1200 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;1,1,0,2,VRT_PF_LEAD=0;1,1,0,2,VRT_PF_LEAD=8;1,1,0,2,VRT_PF_LEAD=16;1,1,0,2,VRT_PF_LEAD=32;1,1,0,2,VRT_PF_LEAD=64" --out gpurun_out/r2i_prefetch_16m.jsonl > gpurun_out/r2i_prefetch_16m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2i_prefetch_16m.jsonl
timeout 600 python profiles/microbench/order_probe.py --workload nlte_1m_native --configs "1,1,0,12;1,1,0,12,VRT_PF_LEAD=0;1,1,0,12,VRT_PF_LEAD=8;1,1,0,12,VRT_PF_LEAD=24;1,1,0,4,VRT_PF_LEAD=8" --out gpurun_out/r2i_prefetch_1m.jsonl > gpurun_out/r2i_prefetch_1m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2i_prefetch_1m.jsonl
