"""Full-size CPU check of the numbers bench.py prints in `results_check` (BASELINE config 2: 1000 draws x 50 trees,
1e6 uniform MC points, 100 candidates): the oracle's C port evaluates the SAME seeded forests and points on the host
cores (about a minute on 8 cores) and the values are compared with a bench line measured on the GPU.

    python tools/bench_expected_cpu.py [profiles/bench_r1_n1.json]          -> prints both, exits 1 on disagreement

bench.py alternates two forests (seeds 2000, 2001); results_check comes from the last timed step, i.e. forest
(steps - 1) % 2.  Test infrastructure (uses oracle/); not product code."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bartint_b200 import synth          # noqa: E402
from oracle import oracle_c as oc       # noqa: E402
from oracle import oracle_np as onp     # noqa: E402

CFG = 2


def expected(forest_index, n_points=None):
    c = dict(synth.CONFIGS[CFG])
    N = n_points or c["N"]
    x_train, y_train = synth.make_design(CFG)
    ff = synth.prior_forest(2000 + forest_index, c["S"], c["m"], x_train, y_train)
    nthreads = max(len(os.sched_getaffinity(0)), 1)
    integ = oc.exact_integrals(ff, "uniform", nthreads=nthreads)
    _, imean, isd = onp.finalize_integrals(integ, ff.y_min, ff.y_max)
    X = synth.make_points(CFG, N, seed=12345)                   # rank 0's points
    dm = oc.predict_reduce(ff, X, nthreads=nthreads, want=("draw_mean",))["draw_mean"]
    Xc = synth.make_points(CFG, c["C"], seed=777)
    var = oc.predict_reduce(ff, Xc, nthreads=nthreads, want=("point_var",))["point_var"]
    return {"integral_mean": float(imean), "integral_sd": float(isd), "mc_mean": float(onp.r_mean(dm)),
            "argmax_candidate": int(np.argmax(var))}


def write_golden():
    """tests/golden/bench_cfg2_expected.json: the values for both forests, read by bench.py (which may not import the
    oracle on its GPU arm) to mark every full-size line as checked."""
    out = {"config": "cfg2 full size: S=1000, m=50, d=10, N=1e6 uniform points (seed 12345), C=100 (seed 777)",
           "made_by": "python tools/bench_expected_cpu.py --write-golden   (oracle/oracle_c.c on the host cores)",
           "forests": {str(2000 + i): expected(i) for i in range(2)}}
    with open(os.path.join(ROOT, "tests", "golden", "bench_cfg2_expected.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


def main():
    if "--write-golden" in sys.argv:
        return write_golden()
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "bench_r1_n1.json")
    j = json.loads(open(path).read().strip().splitlines()[-1])
    got = j["results_check"]
    idx = (j["steps"] - 1) % 2                    # the timed loop restarts its step counter at 0
    n = j["config"]["mc_points_per_gpu"]
    want = expected(idx, n)
    ok = True
    for k, w in want.items():
        g = got[k]
        rel = abs(g - w) / max(abs(w), 1e-300) if isinstance(w, float) else float(g != w)
        print(f"{k:18s} gpu {g!r:24} cpu oracle {w!r:24} rel.diff {rel:.2e}")
        ok &= rel <= 1e-9
    print("forest seed", 2000 + idx, "| points", n, "| AGREE (<= 1e-9 relative; north-star bar 1e-6)" if ok else "| DISAGREE")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
