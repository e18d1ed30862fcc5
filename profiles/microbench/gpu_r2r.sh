#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2r_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
