#!/bin/bash
# GPU call r2a: parity of the lock-step adjoint, bench line, policy variants, ncu capture of the adjoint kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python tools/adj_variants.py > gpurun_out/r2a_variants.log 2>&1
for v in ipol1000 ipol1 u4; do python tools/adj_variants.py tools/_variants/libxde_$v.so >> gpurun_out/r2a_variants.log 2>&1; done
cat gpurun_out/r2a_variants.log
python tools/adj_variants.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dopri5_adj_kernel -s 3 -c 1 -o gpurun_out/r2a_adj python tools/adj_variants.py > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
