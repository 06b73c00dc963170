#!/usr/bin/env python
"""Host topology and NUMA-aware placement of the page-locked point matrix (run with gpurun --gpus N): the aggregate H2D
rate of N concurrent 80 MB shard copies and the one-process group design step, for the matrix allocated with the
default policy, interleaved over the nodes, and with every GPU's shard on that GPU's node.  One JSON line."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bartint_b200 as bi
from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts, synth
sys.path.insert(0, os.path.join(ROOT, "tools"))
import hostmem

G = torch.cuda.device_count()
out = {"gpus": G, "cores": len(os.sched_getaffinity(0)), "numa_nodes": hostmem.numa_nodes(),
       "gpu_numa_node": [hostmem.gpu_numa_node(g) for g in range(G)]}
try:
    out["node_cpus"] = {n: open(f"/sys/devices/system/node/node{n}/cpulist").read().strip() for n in out["numa_nodes"]}
    out["affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["..."] + sorted(os.sched_getaffinity(0))[-2:]
except Exception as e:
    out["node_cpus"] = str(e)
try:
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()[:G + 1]
except Exception as e:
    out["topo"] = str(e)
N, d = 1_000_000, 10
c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
strings, fits = ff.to_dbarts_state(x_train)
wire = WireStrings(strings)
Xc = torch.from_numpy(np.asfortranarray(synth.make_points(2, c["C"], seed=777)).T.copy()).pin_memory().numpy().T
pts = synth.make_points(2, N, seed=1)
dev = [torch.empty((d, N), dtype=torch.float64, device=f"cuda:{g}") for g in range(G)]
streams = [torch.cuda.Stream(device=g) for g in range(G)]
if G > 1:
    bi.init_group(G)
ref = None
for policy in ("default", "interleave", "local"):
    X, info = hostmem.numa_local_pinned_matrix(N, d, out["gpu_numa_node"], policy)
    for g in range(G):
        X[g * N:(g + 1) * N, :] = pts
    r = {"info": info}
    # raw: shard g of every column -> GPU g (d pieces of 8 MB per GPU), all GPUs at once
    Xt = torch.from_numpy(X.T)                     # [d, N_total] row-major view of the same memory
    ts = []
    for rep in range(6):
        for g in range(G):
            torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        for j in range(d):
            for g in range(G):
                with torch.cuda.stream(streams[g]):
                    dev[g][j].copy_(Xt[j, g * N:(g + 1) * N], non_blocking=True)
        for g in range(G):
            streams[g].synchronize()
        ts.append(time.perf_counter() - t0)
    r["h2d_all_shards_ms"] = round(1e3 * float(np.median(ts[1:])), 3)
    r["h2d_GBps_total"] = round(G * N * d * 8 / float(np.median(ts[1:])) / 1e9, 1)
    ts = []
    for rep in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sx, sxc = StagedMatrix(X, "group" if G > 1 else 0), StagedMatrix(Xc, 0)
        o = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0)
        ts.append(time.perf_counter() - t0)
        sx.close(); sxc.close()
    r["group_step_ms"] = round(1e3 * float(np.median(ts[3:])), 3)
    r["group_step_min_ms"] = round(1e3 * float(min(ts[3:])), 3)
    if ref is None:
        ref = o["draw_mean"].copy()
    r["same_result"] = bool(np.array_equal(ref, o["draw_mean"]))
    out[policy] = r
    del X, Xt
if G > 1:
    bi.shutdown_group()
hostmem.release_all()
print(json.dumps(out))
