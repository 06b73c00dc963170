#!/bin/bash
# round 2: the full bench line as the driver runs it (1 GPU), then the contract test
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; echo "bench exit $?"; tail -5 gpurun_out/b_bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; echo "ref exit $?"
timeout 600 python -m pytest tests/test_bench_contract.py -m gpu -x -q 2>&1 | tail -5
