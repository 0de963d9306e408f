#!/bin/bash
# round-2 single-GPU evidence run: every BASELINE workload + kernel microbenches + learn-kernel timings + a launch list
mkdir -p gpurun_out
for w in c3 c1 c2 c4 c5; do
  extra="--no-variants"; [ "$w" = "c3" ] && extra=""
  timeout 600 python bench.py --workload $w $extra > gpurun_out/r02_bench_${w}.json 2> gpurun_out/r02_bench_${w}.err; echo "$w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 300 python scripts/bench_kernels.py > gpurun_out/r02_kernels.jsonl 2> gpurun_out/r02_kernels.err; echo "kernels rc=$?"
timeout 300 python scripts/bench_learn_kernels.py > gpurun_out/r02_learn_kernels.jsonl 2> gpurun_out/r02_learn_kernels.err; echo "learn kernels rc=$?"
timeout 300 python scripts/bench_learn_kernels.py --dropout 0 > gpurun_out/r02_learn_kernels_nodrop.jsonl 2>/dev/null
python scripts/profile_learn.py > gpurun_out/r02_learn_profile_fused.txt 2>&1
python - <<'PY'
import json
for w in ('c3','c1','c2','c4','c5'):
    try:
        r=json.loads(open(f'gpurun_out/r02_bench_{w}.json').read().strip().splitlines()[-1])
        print(w, 'rollout %.3g'%r['value'], 'e2e %.3g'%r['e2e']['value'], 'learn %.3g'%r['learn_value'], 'roof %.3f'%r['roofline']['frac'], 'ms/mb %.3f'%r['learn']['ms_per_minibatch'])
    except Exception as e: print(w, 'ERR', e)
PY
