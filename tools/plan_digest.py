"""Digest of the host planner's output over a fixed set of forests, kernel formats and counting levels (no GPU needed).
Two builds of the library that print the same digests upload byte-identical forests, so a host-side refactor can be
checked against a build whose kernels were verified on a GPU:

    python tools/plan_digest.py [--lib path/to/other/libbartint_cuda.so] > digests.json
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KEYS = ("hdr", "split", "desc", "tree_perm", "group_tree_off", "draw_group_off", "node_bit", "walk_nodes", "leaf_off",
        "leaf_mu")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--levels", default="012")
    a = ap.parse_args()
    from bartint_b200 import _lib, synth
    if a.lib:                       # an older build may lack newer entry points: bind what it has
        import ctypes
        _lib.LIB_PATH = os.path.abspath(a.lib)
        probe = ctypes.CDLL(_lib.LIB_PATH)
        for name in [n for n in _lib.SIGNATURES if not hasattr(probe, n)]:
            del _lib.SIGNATURES[name]
    import plan_interp as pi

    def digest(p):
        h = hashlib.sha256()
        for k in KEYS:
            h.update(np.ascontiguousarray(p[k]).tobytes())
        return [h.hexdigest()[:16], p["n_groups"]]

    res = {}
    for kernel in ("gather", "table"):
        for count in a.levels:
            os.environ["BARTINT_KERNEL"] = kernel
            os.environ["BARTINT_COUNT_SPLITS"] = count
            for d in (1, 2, 5, 8, 10, 12, 13, 16, 20):
                rng = np.random.default_rng(d)
                ff = synth.prior_forest(d, 40, 50, rng.random((100, d)))
                res[f"{kernel}/{count}/d{d}"] = digest(pi.debug_plan(ff))
            for kind in ("stumps", "left_chain", "right_chain", "same_var", "balanced", "edge_cuts"):
                ff = synth.adversarial_forest(kind, 5, 20, 14, seed=3, depth=7)
                res[f"{kernel}/{count}/{kind}"] = digest(pi.debug_plan(ff))
    json.dump(res, sys.stdout, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
