#!/usr/bin/env python
"""Sweep of the draw-chunk count of bartint_design_step_dbarts098 on the benchmark workload (cfg 2): wall clock per
end-to-end step (host buffers in, host results out).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts, synth

c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
strings, fits = ff.to_dbarts_state(x_train)
wire = WireStrings(strings)
X = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["N"])).T.copy()).pin_memory().numpy().T
Xc = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["C"], seed=777)).T.copy()).pin_memory().numpy().T
out = {}
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 12]:
    ts = []
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sx, sxc = StagedMatrix(X), StagedMatrix(Xc)
        o = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0,
                               n_chunks=chunks)
        ts.append(time.perf_counter() - t0)
        sx.close(); sxc.close()
    out[chunks] = round(1e3 * float(np.median(ts[2:])), 3)
print(json.dumps({"ms_per_step_by_chunks": out}))
