#!/bin/bash
# usage: gpu_ncu_sub.sh <bench name filter> <ncu kernel regex> <out name>
F=${1:-ff_bwd}; K=${2:-ff_bwd}; O=${3:-r02_sub_bwd}
mkdir -p gpurun_out
python scripts/bench_learn_kernels.py --only $F --iters 2 > gpurun_out/${O}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -f -o gpurun_out/$O python scripts/bench_learn_kernels.py --only $F --iters 2 > gpurun_out/${O}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${O}_ncu.log
