#!/usr/bin/env python
"""e2e step time against the share of the draws in the first of two chunks (BARTINT_STEP_FIRST), each in a fresh process."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for pct in [int(a) for a in sys.argv[1:]] or [0, 40, 50, 60, 70, 80]:
    env = dict(os.environ, BARTINT_STEP_FIRST=str(pct))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_chunks.py"), "2"], capture_output=True, text=True, env=env)
    out[pct] = json.loads(r.stdout.strip().splitlines()[-1])["ms_per_step_by_chunks"]["2"]
print(json.dumps({"ms_per_step_by_first_chunk_percent (0 = default weights)": out}))
