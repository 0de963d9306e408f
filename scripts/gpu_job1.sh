#!/bin/bash
# round-2 GPU job 1: full GPU test suite, smoke, default bench, kernel microbenches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke1.log
timeout 600 python bench.py --no-variants > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
timeout 300 python scripts/bench_kernels.py > gpurun_out/r2_kernels1.jsonl 2> gpurun_out/r2_kernels1.err; echo "kernels rc=$?"
timeout 300 python scripts/bench_learn_kernels.py > gpurun_out/r2_learn_kernels1.jsonl 2> gpurun_out/r2_learn_kernels1.err; echo "learn kernels rc=$?"
cut -c1-400 gpurun_out/r2_bench1.json
