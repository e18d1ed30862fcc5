#!/bin/bash
# round 2, GPU call L: directions split by wavelength (single-GPU test of the engine)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2l_pytest.log
