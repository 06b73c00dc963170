#!/bin/bash
# ncu launch list of the bench step (no e2e / cpu / strong legs), after the same command ran clean without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-strong"
timeout 120 $CMD > gpurun_out/ll_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "launch list exit $?"
