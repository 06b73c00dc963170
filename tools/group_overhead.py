#!/usr/bin/env python
"""Fixed cost of the device-group path of the design step: the same small workload (cfg-2 forest, few MC points, so the
bus transfer is negligible) through one device and through a group (real GPUs, or BARTINT_GROUP_DEVICES=0,0 on one).
BARTINT_TIMING=1 prints the library's host timestamps to stderr.  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bartint_b200 as bi
from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts, synth

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_each = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
strings, fits = ff.to_dbarts_state(x_train)
wire = WireStrings(strings)
Xc = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["C"], seed=777)).T.copy()).pin_memory().numpy().T
X = torch.from_numpy(np.asfortranarray(synth.make_points(2, n_each * G, seed=3)).T.copy()).pin_memory().numpy().T
out = {"members": G, "rows_each": n_each}


def run(target):
    ts = []
    for rep in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sx, sxc = StagedMatrix(X, target), StagedMatrix(Xc, 0)
        o = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0)
        ts.append(time.perf_counter() - t0)
        sx.close(); sxc.close()
    return round(1e3 * float(np.median(ts[4:])), 3), round(1e3 * float(min(ts[4:])), 3), o["draw_mean"]


# the bare bus transfer of the same shards (d column pieces per GPU, all GPUs at once)
if len(set(os.environ.get("BARTINT_GROUP_DEVICES", "0,1").split(","))) > 1 and torch.cuda.device_count() >= G:
    Xt = torch.from_numpy(X.T)
    devb = [torch.empty((Xt.shape[0], n_each), dtype=torch.float64, device=f"cuda:{g}") for g in range(G)]
    strm = [torch.cuda.Stream(device=g) for g in range(G)]
    ts = []
    for rep in range(5):
        for g in range(G):
            torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        for j in range(Xt.shape[0]):
            for g in range(G):
                with torch.cuda.stream(strm[g]):
                    devb[g][j].copy_(Xt[j, g * n_each:(g + 1) * n_each], non_blocking=True)
        for g in range(G):
            strm[g].synchronize()
        ts.append(time.perf_counter() - t0)
    out["bare_h2d_of_the_shards_ms"] = round(1e3 * float(np.median(ts[1:])), 3)
    del devb
os.environ["BARTINT_GROUP_MIN_ROWS"] = "1"
a = run(0)
out["one_device_ms (median, min)"] = a[:2]
assert bi.init_group(G) == G
b = run("group")
out["group_ms (median, min)"] = b[:2]
out["max_abs_diff"] = float(np.max(np.abs(a[2] - b[2])))
bi.shutdown_group()
print(json.dumps(out))
