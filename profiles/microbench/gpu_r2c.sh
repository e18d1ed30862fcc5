#!/bin/bash
# round 2, GPU call C: new opacity kernel + large parity tests + output file test, kernel variants at 16 M, the other workloads
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c_pytest.log; cat gpurun_out/parity_large.json
timeout 900 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;1,1,0,2,VRT_EXPERIMENT=8;1,1,0,2,VRT_TMA_CFG=26;1,1,0,2,VRT_TMA_CFG=36;1,1,0,2,VRT_TMA_SMEM_KB=75;1,1,0,2,VRT_RUN_LEN=16" --out gpurun_out/r2c_variants_16m.jsonl > gpurun_out/r2c_variants_16m.log 2>&1; echo "variants rc=$?"; cat gpurun_out/r2c_variants_16m.jsonl
timeout 600 python bench.py --workload continuum_1m --steps 10 --warmup 3 > gpurun_out/r2c_bench_continuum_1m.json 2> gpurun_out/r2c_bench_continuum_1m.err; echo "continuum rc=$?"; cat gpurun_out/r2c_bench_continuum_1m.json; tail -n 3 gpurun_out/r2c_bench_continuum_1m.err
timeout 600 python bench.py --workload searchlight --steps 5 --warmup 3 > gpurun_out/r2c_bench_searchlight.json 2> gpurun_out/r2c_bench_searchlight.err; echo "searchlight rc=$?"; cat gpurun_out/r2c_bench_searchlight.json; tail -n 3 gpurun_out/r2c_bench_searchlight.err
