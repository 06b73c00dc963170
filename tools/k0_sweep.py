#!/usr/bin/env python
"""K0 (point binning) bandwidth sweep: the HBM-bound kernel of the path.

For each N the fp64 matrix (N x d column-major) is resident in HBM; the timed region is the points_from_device call
(stream-ordered allocation + the binning launch) measured with CUDA events, L2 flushed before every repetition.
Algorithmic bytes: 8 B read per matrix element + 16 B packed record per point (gather forests) or 1 B feature-major
code per element (table forests).  Prints one JSON line per N.

    python tools/k0_sweep.py [--d 10] [--n 1000000 4000000 12500000 50000000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bartint_b200 import Forest, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--n", type=int, nargs="*", default=[1_000_000, 4_000_000, 12_500_000, 50_000_000])
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    d = args.d
    rng = np.random.default_rng(7)
    x_train = rng.random((20 * d, d))
    ff = synth.prior_forest(7, 8, 10, x_train, None)
    f = Forest.from_soa(ff)
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peak.get("hbm_gbs") or peak.get("hbm", {}).get("gbs") or 0) or None
    except Exception:
        peak = None
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for N in args.n:
        g = torch.Generator(device="cuda").manual_seed(N)
        dX = torch.rand((d, N), dtype=torch.float64, device="cuda", generator=g)
        ms = []
        for _ in range(args.reps + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pts = f.points_from_device(dX.data_ptr(), N, stream=st)
            b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
            pts.close()
        ms = sorted(ms[2:])
        med = ms[len(ms) // 2]
        packed = f.fast_kernel == "gather"
        alg = (8 * N * d + 16 * N) if packed else 9 * N * d      # gather forests: fp64 in, 16 B packed record out
        out = {"N": N, "d": d, "layout": "packed records (feature-major rows on demand)" if packed else "feature-major", "ms_median": med,
               "ms_min": ms[0], "alg_bytes": alg, "gbs": alg / (med * 1e-3) / 1e9,
               "hbm_peak_gbs": peak, "frac": (alg / (med * 1e-3) / 1e9 / peak) if peak else None}
        print(json.dumps(out), flush=True)
        del dX
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
