#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_learn_kernels.py --only ff_bwd --iters 2 > gpurun_out/r02_sub_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ff_bwd" -s 3 -c 1 -f -o gpurun_out/r02_sub_bwd python scripts/bench_learn_kernels.py --only ff_bwd --iters 2 > gpurun_out/r02_sub_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_sub_ncu.log
