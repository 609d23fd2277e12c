#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_adj_tile.py > gpurun_out/r2f_diag.log 2>&1; tail -14 gpurun_out/r2f_diag.log
echo "--- flush every fold"
timeout 600 python tools/diag_adj_tile.py tools/_variants/libxde_flush1.so > gpurun_out/r2f_diag_flush1.log 2>&1; tail -14 gpurun_out/r2f_diag_flush1.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py::test_adjoint_large_state_parity > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2f_pytest.log
