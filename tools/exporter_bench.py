"""Times the HOST phases of forest creation (dbarts-0.9-8 string/fit exporter, validation, walk-form packer, group
planner) through the host-only hook bartint_debug_export_dbarts098 -- no GPU needed.  These phases are what the
sequential-design step hides behind the evaluation kernels ("the step as a pipeline") and what
bounds the end-to-end step when the host threads are divided among 8 ranks.

    python tools/exporter_bench.py [--cfg 2] [--threads 1 2 4 8] [--reps 5]      (BARTINT_TIMING=1 prints the phases)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hosthooks  # noqa: E402   (tests/hosthooks.py: the host-only hooks of include/bartint_test_hooks.h)
from bartint_b200 import WireStrings, _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--threads", type=int, nargs="*", default=[1, 2, 4, 8])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    shapes = {2: (10, 200, 50, 1000), 3: (20, 400, 200, 2000), 5: (10, 40, 50, 1000)}
    d, n, m, S = shapes[a.cfg]
    rng = np.random.default_rng(a.cfg)
    x_train = rng.random((n, d))
    y_train = synth.genz_gaussian(x_train)
    ff = synth.prior_forest(a.cfg, S, m, x_train, y_train)
    strings, fits = ff.to_dbarts_state(x_train)
    wire = WireStrings(strings)
    fits = np.asfortranarray(np.asarray(fits, np.float64).reshape(n, m, S))      # converted once, outside the timing
    xt = np.asfortranarray(x_train)
    run = lambda: hosthooks.export_dbarts(wire, fits, xt, ff.cut_offsets, ff.cut_values, want_arrays=False)[0]
    lib = _lib.load()
    for t in a.threads:
        lib.bartint_set_host_threads(t)
        run()
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            nn = run()
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"cfg": a.cfg, "host_threads": t, "trees": S * m, "nodes": nn,
                          "ms_min": round(1e3 * min(ts), 3), "ms_median": round(1e3 * float(np.median(ts)), 3)}))


if __name__ == "__main__":
    main()
