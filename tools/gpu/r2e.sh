#!/bin/bash
# GPU call r2e: large-state adjoint (tiles + TMEM gradient accumulators), batch adjoint with conflict-free limbs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "large_state or batch" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2e_pytest.log
timeout 300 python tools/bench_configs.py cfg2_batch > gpurun_out/r2e_cfg2_batch.log 2>&1; cut -c1-260 gpurun_out/r2e_cfg2_batch.log
timeout 300 python tools/bench_configs.py cfg3_adjoint > gpurun_out/r2e_cfg3_adjoint.log 2>&1; cut -c1-600 gpurun_out/r2e_cfg3_adjoint.log
