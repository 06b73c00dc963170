#!/usr/bin/env python
"""predict()'s full S x N matrix at the cfg-5 shape (2e5 survey candidates x 10 features, 1000 draws x 50 trees): kernel
time of the evaluation that writes the 1.6 GB matrix (R layout, draws contiguous per point) and the implied store rate."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bartint_b200 import Forest, synth, _lib
import ctypes as C

c = synth.CONFIGS[5]
x_train, y_train = synth.make_design(5)
ff = synth.prior_forest(5001, c["S"], c["m"], x_train, y_train)
N = int(sys.argv[1]) if len(sys.argv) > 1 else c["C"]
X = synth.make_points(5, N)
out = {}
for kern in ("gather", "table"):
    os.environ["BARTINT_KERNEL"] = kern
    f = Forest.from_soa(ff)
    pts = f.points(X)
    full = torch.empty((N, ff.S), dtype=torch.float64, device="cuda")       # column-major S x N
    lib = _lib.load()
    times = []
    for rep in range(4):
        # device-resident full matrix through the internal path: bartint_eval_points would copy 1.6 GB to the host
        rc = lib.bartint_eval_points_full_device(f._h, pts._h, 0, 0, ff.S, C.c_void_p(full.data_ptr()), None)
        assert rc == 0, lib.bartint_last_error()
        torch.cuda.synchronize()
        times.append(f.last_eval_profile(True)["kernel_ms"])
    ms = float(np.median(times[1:]))
    out[kern] = {"kernel_ms": ms, "store_GBps": 8.0 * N * ff.S / (ms * 1e-3) / 1e9, "evals_per_s": ff.S * ff.m * N / (ms * 1e-3)}
    chk = f.eval(X[:64], want=("full_pred",))["full_pred"]
    assert np.allclose(full[:64].cpu().numpy().T, chk, rtol=0, atol=1e-12)
    pts.close(); f.close()
print(json.dumps({"N": N, "S": ff.S, "matrix_GB": 8.0 * N * ff.S / 1e9, **out}))
