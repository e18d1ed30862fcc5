#!/bin/bash
# round 2, GPU call H: final single-GPU evidence
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench_16m.json 2> gpurun_out/r2h_bench_16m.err; echo "bench16m rc=$?"; cat gpurun_out/r2h_bench_16m.json
KREG='regex:k_(sweep_tma|opacity|J_reduce|rates|source_update|stateq_soa|gamma|boundary_planck|boundary_rows|criterion)'
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/r2h_launches_nlte_16m.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_opacity -c 1 -o gpurun_out/r2h_opacity_1m -f python bench.py --workload nlte_1m_native --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_ncu_opacity.log 2>&1; echo "ncu opacity rc=$?"
timeout 900 python bench.py --workload regular_400 --steps 3 --warmup 3 > gpurun_out/r2h_bench_regular_400.json 2> gpurun_out/r2h_bench_regular_400.err; echo "regular rc=$?"; cat gpurun_out/r2h_bench_regular_400.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err; echo "reference rc=$?"; cat gpurun_out/r2h_bench_reference.json; tail -n 3 gpurun_out/r2h_bench_reference.err
