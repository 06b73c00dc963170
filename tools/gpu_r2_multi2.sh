#!/bin/bash
# N-GPU box: the bench as the driver launches it at N = all GPUs and N = 4, then the group probe
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for n in $N 4; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/m2_bench_n$n.json 2> gpurun_out/m2_bench_n$n.err; echo "bench n=$n exit $?"
done
timeout 600 python tools/group_probe.py 2>/dev/null > gpurun_out/m2_group_probe_n$N.json; echo "probe exit $?"; cat gpurun_out/m2_group_probe_n$N.json
