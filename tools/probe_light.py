import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bartint_b200 import Forest, synth
c = synth.CONFIGS[2]
x_train, y_train = synth.make_design(2)
ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
X = synth.make_points(2, c["N"]); Xc = synth.make_points(2, 100, seed=777)
for kern in ("gather", "table"):
    os.environ["BARTINT_KERNEL"] = kern
    f = Forest.from_soa(ff)
    pts = f.points(X); pc = f.points(Xc)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f.eval(pts, want=("draw_mean",)); t1 = time.perf_counter()
        f.eval(pc, want=("point_var",), force_walk=True); t2 = time.perf_counter()
        f.eval(pc, want=("point_var",)); t3 = time.perf_counter()
    print(kern, "mc eval ms", 1e3*(t1-t0), f.last_eval_profile(False), "cand walk ms", 1e3*(t2-t1), "cand fast ms", 1e3*(t3-t2), f.last_eval_profile(True))
    pts.close(); pc.close(); f.close()
