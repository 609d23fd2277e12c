#!/bin/bash
# GPU call r2b: full GPU test suite (incl. round-2 tests), default bench line, e2e breakdown, the cfg3/4/5 lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err
python tools/e2e_breakdown.py 131072 > gpurun_out/r2b_e2e_131072.log 2>&1; tail -12 gpurun_out/r2b_e2e_131072.log
python tools/e2e_breakdown.py 1048576 > gpurun_out/r2b_e2e_1048576.log 2>&1; tail -6 gpurun_out/r2b_e2e_1048576.log
for c in cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r2b_bench_$c.err
  python bench.py --config $c --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_ref_$c.json 2> gpurun_out/r2b_ref_$c.err; echo "$c ref rc=$?"
done
python bench.py --scaling strong --batch 131072 --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2b_bench_131072.json 2> gpurun_out/r2b_bench_131072.err; echo "b17 rc=$?"
