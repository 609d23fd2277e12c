#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_adj_tile.py > gpurun_out/r2g_diag.log 2>&1; tail -14 gpurun_out/r2g_diag.log
timeout 900 python -m pytest tests -m gpu -q -k "large_state" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2g_pytest.log
timeout 300 python tools/bench_configs.py cfg3_adjoint > gpurun_out/r2g_cfg3_adjoint.log 2>&1; cut -c1-600 gpurun_out/r2g_cfg3_adjoint.log
