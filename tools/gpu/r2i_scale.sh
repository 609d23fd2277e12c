#!/bin/bash
# 8-GPU box: strong scaling of cfg2 (the 1M-trajectory batch split N ways) at N = 1, 2, 4, 8 and cfg3 at N = 1, 8
mkdir -p gpurun_out
run() {  # N, extra args, tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 "$@" > gpurun_out/r2i_${tag}_n1.json 2> gpurun_out/r2i_${tag}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n "$@" > gpurun_out/r2i_${tag}_n$n.json 2> gpurun_out/r2i_${tag}_n$n.err
  fi
  echo "$tag N=$n rc=$?"; tail -c 300 gpurun_out/r2i_${tag}_n$n.err | tail -2
}
for n in 1 2 4 8; do run $n strong --scaling strong --steps 20 --warmup 3 --no-cpu --no-secondary; done
for n in 1 8; do run $n cfg3 --config cfg3 --steps 10 --warmup 3; done
run 8 cfg3strong --config cfg3 --scaling strong --steps 10 --warmup 3
run 8 weak --steps 10 --warmup 3 --no-cpu --no-secondary
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        p = json.loads(open(f).read().strip().splitlines()[-1])
        e = p.get('e2e', {})
        print(f.split('/')[-1], 'N', p['n_gpus'], p['scaling'], 'value %.4g' % p['value'], 'ms %.3f' % p['ms_per_step'],
              'e2e %.4g' % e.get('value', 0), 'graph', (e.get('with_cuda_graph') or {}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
