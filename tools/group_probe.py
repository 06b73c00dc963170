#!/usr/bin/env python
"""Where the one-process multi-GPU step spends its time (run with gpurun --gpus N): the platform's concurrent H2D rate
from one process (torch copies, pinned memory, one stream per GPU), then the library's group design step at 1, 2, 4, N
members with BARTINT_TIMING-style wall clocks."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bartint_b200 as bi
from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts, synth

G = torch.cuda.device_count()
out = {"gpus": G, "cores": len(os.sched_getaffinity(0))}
# (a) raw concurrent H2D: 80 MB per GPU from pinned memory, all GPUs at once
N, d = 1_000_000, 10
host = [torch.empty((d, N), dtype=torch.float64).pin_memory() for _ in range(G)]
dev = [torch.empty((d, N), dtype=torch.float64, device=f"cuda:{g}") for g in range(G)]
streams = [torch.cuda.Stream(device=g) for g in range(G)]
for n in sorted({1, 2, 4, G}):
    if n > G:
        continue
    ts = []
    for rep in range(5):
        for g in range(n):
            torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        for g in range(n):
            with torch.cuda.stream(streams[g]):
                dev[g].copy_(host[g], non_blocking=True)
        for g in range(n):
            streams[g].synchronize()
        ts.append(time.perf_counter() - t0)
    out[f"h2d_80MB_x{n}_ms"] = round(1e3 * float(np.median(ts)), 3)
    out[f"h2d_x{n}_GBps_total"] = round(n * 80e6 / float(np.median(ts)) / 1e9, 1)
# one big pinned matrix (what the bench stages): shards of one allocation
big = torch.empty((d, N * G), dtype=torch.float64).pin_memory()
ts = []
for rep in range(5):
    for g in range(G):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for g in range(G):
        with torch.cuda.stream(streams[g]):
            dev[g].copy_(big[:, g * N:(g + 1) * N], non_blocking=True)
    for g in range(G):
        streams[g].synchronize()
    ts.append(time.perf_counter() - t0)
out["h2d_shards_of_one_matrix_ms"] = round(1e3 * float(np.median(ts)), 3)
del host, dev, big
# (b) the library's group step
c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
strings, fits = ff.to_dbarts_state(x_train)
wire = WireStrings(strings)
Xc = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["C"], seed=777)).T.copy()).pin_memory().numpy().T
for n in sorted({1, 2, 4, G}):
    if n > G:
        continue
    X = torch.empty((d, N * n), dtype=torch.float64).pin_memory()
    X[:] = torch.from_numpy(synth.make_points(2, N, seed=1).T).repeat(1, n)
    Xn = X.numpy().T
    if n > 1:
        bi.init_group(n)
    ts, stage = [], []
    for rep in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sx, sxc = StagedMatrix(Xn, "group" if n > 1 else 0), StagedMatrix(Xc, 0)
        t1 = time.perf_counter()
        o = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0)
        ts.append(time.perf_counter() - t0); stage.append(t1 - t0)
        sx.close(); sxc.close()
    out[f"group_step_x{n}_ms"] = round(1e3 * float(np.median(ts[3:])), 3)
    out[f"group_stage_call_x{n}_ms"] = round(1e3 * float(np.median(stage[3:])), 3)
    if n > 1:
        bi.shutdown_group()
print(json.dumps(out))
