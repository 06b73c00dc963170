#!/usr/bin/env python
"""Timeline of ONE end-to-end group design step (cfg 2, 10^6 points per GPU): BARTINT_TIMING=2 makes the library print
wall-clock stamps of GPU progress (host callbacks on the streams); this script prints them relative to the start of the
staging call.  Run with gpurun --gpus N."""
import json, os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("BARTINT_TIMING") != "2":
    env = dict(os.environ, BARTINT_TIMING="2")
    r = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    lines = r.stderr.splitlines()
    start = max(i for i, l in enumerate(lines) if l.startswith("[probe] last step t0"))
    t0 = float(lines[start].split()[-1])
    rows = []
    for l in lines[start + 1:]:
        if l.startswith("[bartint-stamp]"):
            t, rest = l[len("[bartint-stamp] "):].split(" ", 1)
            rows.append((round((float(t) - t0) % 1e5, 3), rest))
    rows.sort()
    for t, rest in rows:
        print(f"{t:9.3f} ms  {rest}")
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-2000:])
    sys.exit(0)
import numpy as np
import torch
import bartint_b200 as bi
from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts, synth

G = torch.cuda.device_count() if len(sys.argv) < 2 else int(sys.argv[1])
N, d = (int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000), 10
c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
strings, fits = ff.to_dbarts_state(x_train)
wire = WireStrings(strings)
Xc = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["C"], seed=777)).T.copy()).pin_memory().numpy().T
X = torch.empty((d, N * G), dtype=torch.float64).pin_memory()
X[:] = torch.from_numpy(synth.make_points(2, N, seed=1).T).repeat(1, G)
Xn = X.numpy().T
if G > 1:
    assert bi.init_group(G) == G
ts = []
for rep in range(7):
    torch.cuda.synchronize()
    for g in range(G):
        torch.cuda.synchronize(g)
    time.sleep(0.05)
    if rep == 6:
        sys.stderr.write(f"[probe] last step t0 {time.monotonic() * 1e3 % 1e5:.3f}\n"); sys.stderr.flush()
    t0 = time.perf_counter()
    sx, sxc = StagedMatrix(Xn, "group" if G > 1 else 0), StagedMatrix(Xc, 0)
    o = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0)
    ts.append(time.perf_counter() - t0)
    sx.close(); sxc.close()
if G > 1:
    bi.shutdown_group()
print(json.dumps({"gpus": G, "step_ms": [round(1e3 * t, 3) for t in ts]}))
