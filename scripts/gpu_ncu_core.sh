#!/bin/bash
# ncu --set full of the attention core kernels at the c3 minibatch shape (the plain run first, as the recipe asks)
mkdir -p gpurun_out
python scripts/bench_learn_kernels.py --only attn_core_fwd --iters 3 > gpurun_out/r02_core_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_core_fwd2 -s 3 -c 1 -f -o gpurun_out/r02_attn_core python scripts/bench_learn_kernels.py --only attn_core_fwd --iters 3 > gpurun_out/r02_core_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_core_ncu.log
