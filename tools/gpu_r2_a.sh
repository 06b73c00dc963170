#!/bin/bash
# round 2, call A: the bit-sliced kernel's own parity tests, then the whole GPU matrix, then the bench (no e2e / cpu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bitslice.py -m gpu -x -q > gpurun_out/a_bitslice.log 2>&1; echo "bitslice tests exit $?"; tail -15 gpurun_out/a_bitslice.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_gpu_all.log 2>&1; echo "all gpu tests exit $?"; tail -5 gpurun_out/a_gpu_all.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench exit $?"; head -c 1500 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
