#!/usr/bin/env python
"""Bus rate of bartint_stage_matrix (four row blocks, each a pitched 2-D copy of d column pieces) against bare 1-D copies
of the same pinned column-major matrix, for growing row counts.  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bartint_b200 import StagedMatrix

d = 10
out = {}
for N in (1_000_000, 8_000_000):
    Xt = torch.empty((d, N), dtype=torch.float64).pin_memory()
    Xt.uniform_()
    X = Xt.numpy().T
    dev = torch.empty((d, N), dtype=torch.float64, device="cuda")
    r = {}
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for j in range(d):
            dev[j].copy_(Xt[j], non_blocking=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    r["bare_1d_ms"] = round(1e3 * float(np.median(ts[1:])), 3)
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dev.copy_(Xt, non_blocking=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    r["bare_one_copy_ms"] = round(1e3 * float(np.median(ts[1:])), 3)
    ts, tc = [], []
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sx = StagedMatrix(X)
        sx.flush()
        t1 = time.perf_counter()
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0); tc.append(t1 - t0)
        sx.close()
    r["staged_ms"] = round(1e3 * float(np.median(ts[2:])), 3)
    r["staged_call_ms"] = round(1e3 * float(np.median(tc[2:])), 3)
    r["GBps_bare_vs_staged_transfer"] = [round(N * d * 8 / r["bare_1d_ms"] / 1e6, 1),
                                         round(N * d * 8 / (r["staged_ms"] - r["staged_call_ms"]) / 1e6, 1)]
    out[N] = r
    del dev, Xt, X
print(json.dumps(out))
