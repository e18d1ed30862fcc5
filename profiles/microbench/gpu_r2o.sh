#!/bin/bash
# round 2, GPU call O: final validation of HEAD (full GPU suite, smoke, short benches)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2o_smoke.log
(time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2o_bench_16m.json 2> gpurun_out/r2o_bench_16m.err); echo "bench rc=$?"; cat gpurun_out/r2o_bench_16m.json | cut -c1-600; grep "\[bench\]" gpurun_out/r2o_bench_16m.err | tail -n 4
