#!/bin/bash
# Multi-GPU parity on a box with >= 2 GPUs (gpurun --gpus 2 / 8): both tests of tests/test_multi_gpu.py, each under
# its own timeout.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/verify_multi_gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -rs > gpurun_out/verify_multi_gpu.log 2>&1; echo "multi-gpu exit $?"; tail -5 gpurun_out/verify_multi_gpu.log
