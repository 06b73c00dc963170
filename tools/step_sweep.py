#!/usr/bin/env python
"""e2e step time against (draw chunks, row blocks of X put on the bus up front), each setting in a fresh process."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for up in (1, 2, 3, 4):
    for ch in (1, 2):
        env = dict(os.environ, BARTINT_STAGE_UPFRONT=str(up))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_chunks.py"), str(ch)], capture_output=True, text=True, env=env)
        out[f"upfront{up}_chunks{ch}"] = json.loads(r.stdout.strip().splitlines()[-1])["ms_per_step_by_chunks"][str(ch)]
print(json.dumps(out))
