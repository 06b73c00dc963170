#!/bin/bash
# One gpurun call that verifies, on a B200, everything that was changed after the round-1 GPU budget was spent
# (exporter leaf search, planner refactor, counting level 2), each step under its own timeout so that a hang cannot
# take the box down with it (the first version of level 2 did exactly that):
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/verify_on_gpu.sh'
#
# Results land in gpurun_out/verify_*.log / .jsonl.
mkdir -p gpurun_out
echo "== 1. default GPU matrix"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/verify_gpu_default.log 2>&1; echo "exit $?"; tail -3 gpurun_out/verify_gpu_default.log
echo "== 2. counting level 2 under every GPU parity test"
BARTINT_TEST_COUNT2=1 timeout 900 python -m pytest tests -m gpu -x -q -k count2 > gpurun_out/verify_gpu_count2.log 2>&1; echo "exit $?"; tail -3 gpurun_out/verify_gpu_count2.log
echo "== 2b. the R shim end to end through the mock R API"
BARTINT_TEST_RSHIM_GPU=1 timeout 300 python -m pytest tests/test_r_shim_mock.py -x -q > gpurun_out/verify_rshim.log 2>&1; echo "exit $?"; tail -3 gpurun_out/verify_rshim.log
echo "== 3. Monte Carlo step at counting levels 0 / 1 / 2 (same box)"
: > gpurun_out/verify_count_levels.jsonl
for L in 0 1 2; do
  BARTINT_COUNT_SPLITS=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu >> gpurun_out/verify_count_levels.jsonl 2> gpurun_out/verify_count_$L.err; echo "level $L exit $?"
done
python - <<'PY'
import json
for l in open("gpurun_out/verify_count_levels.jsonl"):
    j = json.loads(l)
    print(j["config"].get("count_splits"), "ms/step", round(j["ms_per_step"], 3), "MC kernel ms", round(j["roofline"]["kernel_ms"], 3),
          "mc_mean", j["results_check"]["mc_mean"])
PY
echo "== 4. the bench as the driver runs it"
timeout 600 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "exit $?"; cat gpurun_out/verify_bench.json | head -c 600; echo
