#!/usr/bin/env python
"""Runs the BASELINE.json configurations other than the benchmark's (cfg 3, 4, 5; cfg 1 as a smoke) on one GPU and
prints one JSON line per config: throughput of the fused evaluation, size-independent parity properties
(table kernel == walk kernel on a sub-sample, MC mean within 5 sigma of the exact integral), and timings.
Not a bench line -- bench.py is; these are the parity/measurement cases of SURVEY.md section 8d.

    python tools/run_configs.py [--cfg 3 4 5] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bartint_b200 import Forest, synth  # noqa: E402


def device_points(cfg, n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    measure = synth.CONFIGS[cfg]["measure"]
    u = torch.rand((d, n), dtype=torch.float64, device="cuda", generator=g)      # [d, n] == n x d column-major
    if measure == "gaussian":      # N(0.5, 1) truncated to [0, 1]: inverse CDF
        lo, hi = 0.3085375387259869, 0.6914624612740131
        u = 0.5 + torch.special.ndtri(lo + (hi - lo) * u)
    elif measure == "survey":
        u = torch.floor(u * torch.tensor([6, 2, 2, 2, 2, 9, 24, 2, 2, 2], dtype=torch.float64, device="cuda")[:, None]) + 1.0
    return u


def run(cfg, scale):
    c = synth.CONFIGS[cfg]
    S, m, d = c["S"], c["m"], c["d"]
    N = int((c["N"] if c["N"] else c["C"]) * scale)
    x_train, y_train = synth.make_design(cfg)
    t0 = time.perf_counter()
    ff = synth.prior_forest(1000 * cfg, S, m, x_train, y_train)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    f = Forest.from_soa(ff)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    dX = device_points(cfg, N, d, cfg)
    st = torch.cuda.current_stream().cuda_stream
    want_mom = c["integrand"] == "survey"
    sums = torch.empty(S, dtype=torch.float64, device="cuda")
    mean = torch.empty(N, dtype=torch.float64, device="cuda") if want_mom else None
    m2 = torch.empty(N, dtype=torch.float64, device="cuda") if want_mom else None
    times, kms = [], []
    for it in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pts = f.points_from_device(dX.data_ptr(), N, stream=st)
        f.eval_device(pts, d_draw_sum=sums.data_ptr(), d_point_mean=mean.data_ptr() if want_mom else 0,
                      d_point_m2=m2.data_ptr() if want_mom else 0, stream=st)
        b.record(); b.synchronize()
        times.append(a.elapsed_time(b)); kms.append(f.last_eval_profile(want_mom)["kernel_ms"])
        if it < 2:
            pts.close()
    prof = f.last_eval_profile(want_mom)
    # parity properties
    n_sub = min(N, 20000)
    sub = f.points_from_device(dX.data_ptr(), n_sub, ld=N, stream=st)
    s_tab = torch.empty(S, dtype=torch.float64, device="cuda"); s_walk = torch.empty_like(s_tab)
    f.eval_device(sub, d_draw_sum=s_tab.data_ptr(), stream=st)
    f.eval_device(sub, d_draw_sum=s_walk.data_ptr(), stream=st, force_walk=True)
    torch.cuda.synchronize()
    rel = float(((s_tab - s_walk).abs() / s_walk.abs().clamp_min(1e-3 * n_sub)).max())
    out = {"cfg": cfg, "workload": c["name"], "S": S, "m": m, "d": d, "points": N, "tree_evals": S * m * N,
           "ms": min(times), "kernel_ms": min(kms), "evals_per_s": S * m * N / (min(times) * 1e-3),
           "kernel": prof.get("kernel"), "table_kernel": prof["table_kernel"], "moments": want_mom, "max_internal_nodes": f.max_internal,
           "groups": f.n_groups, "table_bits": f.table_bits, "table_vs_walk_rel_err": rel,
           "forest_gen_s": t_gen, "forest_upload_s": t_upload}
    if c["measure"] == "uniform":
        t0 = time.perf_counter()
        exact = f.exact_integrals("uniform")
        out["exact_integral_ms"] = 1e3 * (time.perf_counter() - t0)
        mc = sums.cpu().numpy() / N
        # per-draw sd of f over points estimated on the sub-sample
        full = f.eval(dX[:, :n_sub].T.cpu().numpy(), want=("full_pred",), scaled=True)["full_pred"]
        se = full.std(axis=1, ddof=1) / np.sqrt(N)
        out["mc_vs_exact_max_sigmas"] = float((np.abs(mc - exact) / np.maximum(se, 1e-300)).max())
    pts.close(); sub.close(); f.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, nargs="+", default=[3, 5, 4])
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    for cfg in a.cfg:
        run(cfg, a.scale)
