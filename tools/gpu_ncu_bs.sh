#!/bin/bash
# ncu of the bit-sliced kernel at cfg 2 (bench step, no e2e / cpu legs): launch list + one full capture
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bs.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list exit $?"
timeout 300 $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_bitslice -s 1 -c 2 -o gpurun_out/prof_bs -f $CMD > gpurun_out/ncu_f.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_f.log
