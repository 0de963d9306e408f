#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_kernels.py --only loss --iters 2 > gpurun_out/r02_loss_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ppo_loss_rows -s 6 -c 1 -f -o gpurun_out/r02_ppo_loss python scripts/bench_kernels.py --only loss --iters 2 > gpurun_out/r02_loss_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_loss_ncu.log
