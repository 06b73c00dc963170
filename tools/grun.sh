#!/bin/bash
# build locally (the .so travels with the snapshot), then run a command on the GPU box:  tools/grun.sh [--gpus N] TIMEOUT 'cmd'
set -e
cd "$(dirname "$0")/.."
python bart-int_b200/build.py >/dev/null 2>gpurun_out/build.err || { tail -20 gpurun_out/build.err; exit 1; }
make -C oracle -s
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
