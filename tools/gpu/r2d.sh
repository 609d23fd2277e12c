#!/bin/bash
# GPU call r2d: sequence-exact batch-controller adjoint (exact 128-bit batch sums)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "batch or reference_defaults" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2d_pytest.log
python tools/bench_configs.py cfg2_batch > gpurun_out/r2d_cfg2_batch.log 2>&1; cat gpurun_out/r2d_cfg2_batch.log | cut -c1-400
