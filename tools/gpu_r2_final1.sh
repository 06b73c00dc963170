#!/bin/bash
# round 2, 1-GPU box: the whole GPU test matrix, smoke, and both bench arms exactly as the driver runs them
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/f_gpu_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/f_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref exit $?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/f_bench_n1.err
