#!/bin/bash
# one full ncu capture of the gather kernel's predict()-matrix instantiation at the cfg-5 shape
mkdir -p gpurun_out
CMD="python tools/full_pred_probe.py"
timeout 300 $CMD > gpurun_out/full_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_gather -s 2 -c 1 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/full_plain.log
