#!/bin/bash
# round 2, GPU call J: validation of the cleaned tree (full GPU suite, smoke) and the opacity kernel with constant-bank coefficients
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2j_smoke.log
timeout 900 python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2" --iters 3 --out gpurun_out/r2j_probe_16m.jsonl > gpurun_out/r2j_probe_16m.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2j_probe_16m.jsonl
timeout 600 python bench.py --workload nlte_1m_native --steps 5 --warmup 3 > gpurun_out/r2j_bench_1m.json 2> gpurun_out/r2j_bench_1m.err; echo "bench1m rc=$?"; cat gpurun_out/r2j_bench_1m.json | cut -c1-1500; tail -n 4 gpurun_out/r2j_bench_1m.err
