#!/bin/bash
# last call of the round: the whole GPU test matrix, smoke and the bench's own arm (the reference arm was run earlier)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/f_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -2 gpurun_out/f_gpu_tests.log
timeout 100 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/f_smoke.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/f_bench_n1_last.json 2> gpurun_out/f_bench_n1_last.err; echo "bench exit $?"
