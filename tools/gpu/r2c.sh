#!/bin/bash
# GPU call r2c: CUDA-graph e2e at 2^20 and 2^17, host-enqueue diagnostics of the cfg3/4/5 lines
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
python bench.py --scaling strong --batch 131072 --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2c_bench_131072.json 2> gpurun_out/r2c_bench_131072.err; echo "b17 rc=$?"; tail -3 gpurun_out/r2c_bench_131072.err
for c in cfg3 cfg5; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2c_bench_$c.json 2> gpurun_out/r2c_bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r2c_bench_$c.err
done
python - <<'PY'
import json
for f in ['r2c_bench','r2c_bench_131072']:
    p=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, p['ms_per_step'], p['e2e']['ms_per_step'], p['e2e']['with_cuda_graph'])
for c in ['cfg3','cfg5']:
    p=json.loads(open(f'gpurun_out/r2c_bench_{c}.json').read().strip().splitlines()[-1])
    print(c, p['ms_per_step'], p['kernel_ms'], p['e2e']['ms_per_step'], p['config'].get('host_enqueue_ms_per_step'))
PY
