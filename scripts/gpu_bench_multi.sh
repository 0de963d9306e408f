#!/bin/bash
# usage: gpu_bench_multi.sh N workload [workload ...]   -- bench.py under torchrun on N GPUs, one JSON per workload
N=$1; shift
mkdir -p gpurun_out
port=29600
for w in "$@"; do
  port=$((port+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload $w --no-variants --no-cpu > gpurun_out/r02_bench_${w}_n${N}.json 2> gpurun_out/r02_bench_${w}_n${N}.err
  echo "$w N=$N rc=$?"
  python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/r02_bench_${w}_n${N}.json').read().strip().splitlines()[-1])
    print('${w}', 'n', r['n_gpus'], 'rollout %.4g'%r['value'], 'e2e %.4g'%r['e2e']['value'], 'learn %.4g'%r['learn_value'], 'ms/mb %.3f'%r['learn']['ms_per_minibatch'], 'identical', r.get('replicas_identical'), 'coll', r['learn'].get('collective','')[:5], 'opt_share %.4f'%r['learn']['kernel_share']['optimizer'], 'evolve', r.get('evolve',{}).get('ms_per_fitness_gather'))
except Exception as e: print('${w}', 'ERR', e)
PY
done
