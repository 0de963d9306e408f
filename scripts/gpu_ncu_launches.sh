#!/bin/bash
# launch list of one short bench.py run (rollout passes + Agent.learn calls): first plain, then under ncu with one metric
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-variants --no-cpu --no-e2e --learn-steps 1 --learn-epochs 1"
timeout 300 python bench.py $ARGS > gpurun_out/r02_launches_plain.json 2> gpurun_out/r02_launches_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r02_launches_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py $ARGS > gpurun_out/r02_launches_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_launches_c3.csv
