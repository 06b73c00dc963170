#!/bin/bash
# one full ncu capture of the bit-sliced kernel at cfg 2 (no strong lines, no launch list)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-strong"
timeout 300 $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_bitslice -s 1 -c 1 -o gpurun_out/prof_bs3 -f $CMD > gpurun_out/ncu_f3.log 2>&1
echo "full capture exit $?"
