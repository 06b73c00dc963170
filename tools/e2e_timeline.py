#!/usr/bin/env python
"""Where the end-to-end design step (cfg 2, host buffers) spends its wall clock: the bus transfer of the point matrix
alone, the step without Monte Carlo points (export + forest build + exact integrals + candidates), the whole step, and
the library's own per-chunk host timestamps (BARTINT_TIMING=1).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BARTINT_TIMING"] = "1"
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


def med(fn, reps=8, skip=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return round(1e3 * float(np.median(ts[skip:])), 3), round(1e3 * float(min(ts[skip:])), 3)


def h2d_only():
    sx = StagedMatrix(X)
    sx.flush()
    torch.cuda.synchronize()
    sx.close()


def step(mc=True):
    sx = StagedMatrix(X) if mc else None
    sxc = StagedMatrix(Xc)
    design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0)
    if sx is not None:
        sx.close()
    sxc.close()


out = {"h2d_of_X_only_ms (median, min)": med(h2d_only),
       "step_without_mc_points_ms": med(lambda: step(False)),
       "step_ms": med(step)}
print(json.dumps(out))
