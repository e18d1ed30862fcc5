#!/bin/bash
# round 2, GPU call F: directions in flight paired by sense and inclination, proportional pacing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1200 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;1,1,0,2,VRT_PACING=1;1,1,0,2,VRT_DIR_ORDER=1;1,1,0,2,VRT_DIR_ORDER=1,VRT_PACING=1" --out gpurun_out/r2f_pairing_16m.jsonl > gpurun_out/r2f_pairing_16m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2f_pairing_16m.jsonl
timeout 600 python profiles/microbench/order_probe.py --workload nlte_1m_native --configs "1,1,0,12;1,1,0,12,VRT_PACING=1;1,1,0,12,VRT_DIR_ORDER=1,VRT_PACING=1;1,1,0,4,VRT_DIR_ORDER=1,VRT_PACING=1;1,1,0,6,VRT_DIR_ORDER=1,VRT_PACING=1" --out gpurun_out/r2f_pairing_1m.jsonl > gpurun_out/r2f_pairing_1m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2f_pairing_1m.jsonl
