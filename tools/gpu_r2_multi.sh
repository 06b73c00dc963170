#!/bin/bash
# round 2, N-GPU box: group + torch.distributed parity tests, then the bench as the driver launches it (N = all GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_group.py tests/test_multi_gpu.py -m gpu -x -q -rs > gpurun_out/m_tests_n$N.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/m_tests_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err; echo "bench exit $?"; tail -3 gpurun_out/m_bench_n$N.err
