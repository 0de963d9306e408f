#!/bin/bash
# ncu --set full of the step kernel's launch at t=250 of a c3 rollout (the launch whose algorithmic bytes equal the per-launch
# average): the plain run first, as the recipe asks
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-variants --no-cpu --no-e2e --no-learn"
timeout 200 python bench.py $ARGS > gpurun_out/r02_step_plain.json 2> gpurun_out/r02_step_plain.err || { echo "plain run failed"; tail -3 gpurun_out/r02_step_plain.err; exit 1; }
timeout 300 ncu --set full --clock-control none -k regex:wm_step_kernel -s 250 -c 1 -f -o gpurun_out/r02_wm_step_t250 python bench.py $ARGS > gpurun_out/r02_step_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_step_ncu.log
