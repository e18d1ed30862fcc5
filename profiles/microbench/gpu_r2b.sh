#!/bin/bash
# round 2, GPU call B: regression tests, 1 M bench + ncu source-level capture of the sweep kernel, full default bench at 16 M
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --workload nlte_1m_native --steps 5 --warmup 3 > gpurun_out/r2b_bench_1m.json 2> gpurun_out/r2b_bench_1m.err; echo "bench1m rc=$?"; cat gpurun_out/r2b_bench_1m.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_sweep_tma -c 1 -o gpurun_out/r2b_sweep_1m -f python bench.py --workload nlte_1m_native --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_16m.json 2> gpurun_out/r2b_bench_16m.err; echo "bench16m rc=$?"; cat gpurun_out/r2b_bench_16m.json; tail -n 5 gpurun_out/r2b_bench_16m.err
