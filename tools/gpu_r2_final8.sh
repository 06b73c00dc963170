#!/bin/bash
# N-GPU box, end of round 2: group parity tests, the bench as the driver launches it at N = all GPUs, host timestamps of one group step
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_group.py -m gpu -x -q -rs > gpurun_out/f8_tests_n$N.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/f8_tests_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/f8_bench_n$N.json 2> gpurun_out/f8_bench_n$N.err; echo "bench exit $?"; tail -2 gpurun_out/f8_bench_n$N.err
BARTINT_TIMING=1 timeout 200 python tools/group_overhead.py $N 1000000 > gpurun_out/f8_group_n$N.json 2> gpurun_out/f8_group_n$N.err; echo "group exit $?"; cat gpurun_out/f8_group_n$N.json; grep "design step\|members\|chunk 0" gpurun_out/f8_group_n$N.err | tail -6
