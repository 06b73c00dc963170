#!/usr/bin/env python
"""Wall-clock breakdown of one end-to-end sequential-design step through the host-buffer C ABI (cfg 2)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bartint_b200 import Forest, WireStrings, synth

c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
strings, fits = ff.to_dbarts_state(x_train)
wire = WireStrings(strings)
X = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["N"])).T.copy()).pin_memory().numpy().T
Xc = synth.make_points(2, c["C"], seed=777)
out = {}
for rep in range(4):
    t = [time.perf_counter()]
    f = Forest.from_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max); torch.cuda.synchronize(); t.append(time.perf_counter())
    integ = f.exact_integrals("uniform"); t.append(time.perf_counter())
    var = f.eval(Xc, weights=np.ones(1), want=("point_var",)); t.append(time.perf_counter())
    pts = f.points(X); t.append(time.perf_counter())
    dm = f.eval(pts, want=("draw_mean",)); t.append(time.perf_counter())
    pts.close(); t.append(time.perf_counter())
    dm2 = f.eval(X, want=("draw_mean",)); t.append(time.perf_counter())      # one-shot bartint_eval (row-block pipeline)
    f.close(); t.append(time.perf_counter())
    names = ["from_dbarts", "exact_integrals", "candidates_eval", "points(H2D+bin)", "mc_eval", "points_close",
             "one_shot_eval(pipelined H2D+bin+eval)", "close"]
    out = {n: round(1e3 * (b - a), 3) for n, a, b in zip(names, t, t[1:])}
    out["total_two_step_ms"] = round(1e3 * (t[6] - t[0]), 3)
print(json.dumps(out))
