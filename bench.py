#!/usr/bin/env python
"""Benchmark of the BART-Int posterior-evaluation hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- Genz Gaussian-peak d=10, 50 trees x 1000 draws,
1e6 uniform MC points (+ the C=100 sequential-design candidates).  One *step* is one sequential-design
pass of the hot path over a fresh posterior that is already resident in HBM:
    exact leaf-volume integrals of all S draws + their mean/sd        (sampleIntegrals, BARTInt.R:259-262)
    candidates: bin -> ensemble evaluation -> colVars * weights       (BARTInt.R:312-314)
    MC points : bin -> ensemble evaluation -> rowMeans -> mean/var    (bartMean.R:74-82 estimator)
metric = tree-evals/s = S*m*(N + C) per step / device time.  N > 1 ranks: weak scaling, the MC points are
sharded (1e6 per GPU, all-gathered per-draw sums) and the candidate evaluation is sharded over draws
(all-gathered moments, Chan merge).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "tree-evals/sec (draws x trees x points)"
UNIT = "tree-evals/s"
CFG = 2


def workload(args):
    from bartint_b200 import synth
    c = dict(synth.CONFIGS[CFG])
    if args.points:
        c["N"] = args.points
    if args.draws:
        c["S"] = args.draws
    return c


def config_json(c, extra=None):
    out = {"workload": c["name"], "d": c["d"], "trees": c["m"], "draws": c["S"], "mc_points_per_gpu": c["N"],
           "candidates": c["C"], "measure": c["measure"], "cutpoints_per_variable": 100,
           "l2": "flushed between steps (512 MiB memset outside the per-step event pairs)"}
    out.update(extra or {})
    return out


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path = the C restatement (oracle port),
# all host threads, a bounded sample of the same workload per step
# ---------------------------------------------------------------------------------------------
def cpu_sample_run(ff, X, nthreads, budget_s, oc):
    """Time the CPU port on as many of X's rows as fit the time budget.  Returns (evals/s, rows used, seconds)."""
    probe = min(len(X), 2000)
    t0 = time.perf_counter()
    oc.predict_reduce(ff, X[:probe], nthreads=nthreads, want=("draw_mean",))
    rate = ff.S * ff.m * probe / max(time.perf_counter() - t0, 1e-9)
    rows = int(min(len(X), max(probe, budget_s * rate / (ff.S * ff.m))))
    t0 = time.perf_counter()
    oc.predict_reduce(ff, X[:rows], nthreads=nthreads, want=("draw_mean",))
    dt = time.perf_counter() - t0
    return ff.S * ff.m * rows / dt, rows, dt


def host_cores(oc):
    """Threads for the CPU arm: every core this process may run on.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would time the reference on ONE core at N > 1 (only rank 0 runs it, so
    the whole host is its own)."""
    try:
        return max(len(os.sched_getaffinity(0)), oc.max_threads())
    except AttributeError:
        return max(os.cpu_count() or 1, oc.max_threads())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bartint_b200 import synth
    from oracle import oracle_c as oc
    c = workload(args)
    x_train, y_train = synth.make_design(CFG)
    ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
    nthreads = host_cores(oc)
    rows = min(c["N"], 20000 * max(1, nthreads // 4))
    X = synth.make_points(CFG, rows)
    Xc = synth.make_points(CFG, c["C"], seed=777)

    def step():
        oc.exact_integrals(ff, "uniform", nthreads=nthreads)
        oc.predict_reduce(ff, Xc, nthreads=nthreads, want=("point_var",))
        oc.predict_reduce(ff, X, nthreads=nthreads, want=("draw_mean",))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    evals = ff.S * ff.m * (rows + c["C"]) * args.steps
    value = evals / dt
    sample = f"{rows} of {c['N']} MC points + {c['C']} candidates, all {c['S']} draws, per step"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_json(c, {"reference_note": "R/dbarts cannot run here (no R); this is the C restatement of "
                                  "the dbarts predict loop + sampleIntegrals arithmetic (oracle/oracle_c.c), OpenMP over draws"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = sorted(sm)[len(sm) // 2:] if sm else []       # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from bartint_b200 import Forest, synth, _lib
    from bartint_b200 import dist as bdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    # torchrun pins OMP_NUM_THREADS=1 per rank, which would make the exporter (host C++, OpenMP) serial: give every
    # rank its share of the host cores instead
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = int(lib.bartint_set_host_threads(max(1, (os.cpu_count() or 1) // max(local_world, 1))))
    c = workload(args)
    S, m, d, N, C = c["S"], c["m"], c["d"], c["N"], c["C"]
    n_total = N * world

    # ---- inputs (synthetic, seeded): posteriors of two consecutive design steps, MC points, candidates ----
    x_train, y_train = synth.make_design(CFG)
    flats = [synth.prior_forest(2000 + it, S, m, x_train, y_train) for it in range(2)]
    forests = [Forest.from_soa(ff) for ff in flats]
    assert all(f.table_kernel_ok for f in forests), "table kernel must apply to the benchmark forest"
    X_host = torch.from_numpy(np.asfortranarray(synth.make_points(CFG, N, seed=12345 + rank)).T.copy()).pin_memory()  # [d, N]
    Xc_host = torch.from_numpy(np.asfortranarray(synth.make_points(CFG, C, seed=777)).T.copy()).pin_memory()
    dX, dXc = X_host.to(dev), Xc_host.to(dev)
    st = torch.cuda.current_stream().cuda_stream
    d_int = torch.empty(S, dtype=torch.float64, device=dev)
    d_int_res = torch.empty(S, dtype=torch.float64, device=dev)
    d_int_ms = torch.empty(2, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    results = {}

    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    def step(i):
        f = forests[i % len(forests)]
        # the large MC evaluation is enqueued first; the small chains (exact integral; candidates) follow on a side
        # stream so neither their kernels nor their host-side launch work sit in front of it
        side.wait_stream(main)
        pm = f.points_from_device(dX.data_ptr(), N, stream=st)
        dm, mv = bdist.point_sharded_draw_means(f, pm, stream=st, n_total=n_total)                                  # bartMean.R:74-82
        pm.close()
        with torch.cuda.stream(side):
            f.exact_integrals_device(d_int.data_ptr(), "uniform", d_rescaled=d_int_res.data_ptr(),
                                     d_mean_sd=d_int_ms.data_ptr(), stream=side.cuda_stream)    # sampleIntegrals + :260-262
            pc = f.points_from_device(dXc.data_ptr(), C, stream=side.cuda_stream)
            var = bdist.draw_sharded_candidate_var(f, pc, None)                                  # :312-314
            pc.close()
        main.wait_stream(side)
        results.update(var=var, draw_mean=dm, mean_var=mv, forest=f)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # samples from warm-up to the end of the timed region
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    launches0 = lib.bartint_kernel_launch_count()
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.zero_()                                   # evict the previous step's data from the 126 MB L2
        ev[i][0].record()
        step(i)
        ev[i][1].record()
        ev[i][1].synchronize()
        kernel_ms.append(results["forest"].last_eval_profile()["kernel_ms"])    # the MC-point evaluation kernel
    torch.cuda.synchronize()
    launches = lib.bartint_kernel_launch_count() - launches0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    evals_per_step = S * m * (n_total + C)
    value = evals_per_step * args.steps / (total_ms * 1e-3)

    # ---- K0 (binning) on its own: the HBM-bound kernel of the path ----
    k0_ms = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        p = forests[0].points_from_device(dX.data_ptr(), N, stream=st)
        b.record(); b.synchronize()
        k0_ms.append(a.elapsed_time(b))
        p.close()

    # ---- end to end through the reference-facing call: dbarts wire format + host matrices in, host results out ----
    e2e = None
    if not args.no_e2e:
        from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts
        strings, fits = flats[0].to_dbarts_state(x_train)
        strings = WireStrings(strings)      # R holds savedTrees as C strings already; marshal once, outside the timing
        cut_lists = flats[0].cut_lists()
        Xn = X_host.numpy().T          # N x d view, column-major (pinned)
        Xcn = Xc_host.numpy().T
        e2e_steps = max(1, min(args.steps, 3))
        times = []
        h2d = d2h = 0
        for k in range(e2e_steps + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            # the point matrices do not depend on the fitted model: their H2D copies are started first and overlap
            # the exporter's CPU work (bartint_stage_matrix); everything goes through the host-buffer C ABI
            sx, sxc = StagedMatrix(Xn, local_rank), StagedMatrix(Xcn, local_rank)
            if args.e2e_separate:        # the same step as separate ABI calls (exporter, then the evaluations)
                f = Forest.from_dbarts(strings, fits, x_train, cut_lists, flats[0].y_min, flats[0].y_max)   # exporter + H2D
                integ = f.exact_integrals("uniform")
                _, imean, isd = f.finalize_integrals(integ)
                pc, pm = sxc.points(f), sx.points(f)
                var = f.eval(pc, weights=np.ones(1), want=("point_var",))["point_var"]
                dmean = f.eval(pm, want=("draw_mean",))["draw_mean"]
                torch.cuda.synchronize()
                pc.close(); pm.close(); f.close()
            elif world >= 8:             # every rank exports 1/world of the draws; sub-forest images are all-gathered
                # (measured: 11.4 vs 13.5 ms at 8 ranks with 4 host threads each; with more host threads per rank the
                #  per-rank pipelined call below is faster: 7.9 vs 9.3 ms at 2 ranks, 9.5 vs 9.9 ms at 4)
                o = bdist.sharded_design_step(strings, fits, x_train, cut_lists, flats[0].y_min, flats[0].y_max, sx, sxc,
                                              weights=1.0, measure="uniform", n_total=n_total)
                integ, var, dmean = o["integrals"], o["cand_var"], o["draw_mean"]
            else:                        # one call: bartint_design_step_dbarts098 (exporter pipelined against the kernels)
                o = design_step_dbarts(strings, fits, x_train, cut_lists, flats[0].y_min, flats[0].y_max, mc=sx,
                                       candidates=sxc, weights=1.0, measure="uniform")
                integ, var, dmean = o["integrals"], o["cand_var"], o["draw_mean"]
            dt = time.perf_counter() - t0
            sx.close(); sxc.close()
            if k > 0:
                times.append(dt)
            f0 = forests[0]
            h2d = Xn.nbytes + Xcn.nbytes + f0.n_nodes * 9 + f0.n_leaves * 8 + f0.n_groups * 8 * (f0.table_bits + 1) + S * m * 8
            d2h = integ.nbytes * 2 + var.nbytes + dmean.nbytes + 32
        t_e2e = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": S * m * (n_total + C) / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "s_per_step": float(t_e2e.item()), "steps": e2e_steps,
               "call": "separate ABI calls" if args.e2e_separate else
                       ("dist.sharded_design_step (draw-sharded export, all-gathered forest images, point-sharded evaluation)"
                        if world >= 8 else "bartint_design_step_dbarts098 (one call per rank, draw-chunk pipeline)"),
               "includes": "X / candidate H2D from pinned memory (started first, overlapping the exporter), dbarts-0.9-8 "
                           "string/fit exporter, forest H2D + table build, binning, exact integrals, evaluation, D2H of "
                           "integrals, candidate variances and per-draw means"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (MC-point ensemble evaluation) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    k1_ms = float(np.mean(kernel_ms))
    alg_bytes = N * d * 1 + S * m * 26 + 8 * S            # uint8 codes + compact forest + per-draw sums
    evals_k1 = S * m * N
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    # on-chip ceilings of the evaluation kernel: per warp and group, a warp covers 32 * P
    # (group, point) pairs.  Shared-memory pipe: 1 wavefront / clk / SM.  Issue: 4 warp-instructions / clk / SM.
    f0 = forests[0]
    trees_per_group = S * m / max(f0.n_groups, 1)
    nb = f0.table_bits
    if f0.fast_kernel == "gather":
        pts = 16                                         # points per thread (sums-only kernel of the MC path)
        # DUAL record: two 16-entry fp64 tables per record and point
        wavefronts = pts * 2 * 2.0 + 2.0                 # 2 table LDS.64 per point (2 wavefronts each, conflict-free) + 2 descriptor loads
        instrs = pts * 12.0 + 2.0 + 7.0                  # (PRMT, IADD, PRMT, IDP, LDS, DADD) per table and point; descriptors; loop
    else:
        pts = 16
        wavefronts = nb / 2.0 + nb * 4.0 + pts * 2.0     # descriptor LDS.128 + code LDS.32 per (node, quad) + table LDS.64
        instrs = nb * (2.0 + 4 * 4.0) + pts * 3.0 + nb / 2.0 + 5.0
    per_warp = 32 * pts * trees_per_group                # tree-evals per warp-group
    lsu_bound = 148 * sm_mhz * 1e6 * per_warp / wavefronts
    issue_bound = 148 * sm_mhz * 1e6 * per_warp / (instrs / 4.0)
    onchip_bound = min(lsu_bound, issue_bound)
    roofline = {
        "kernel": f"eval_{f0.fast_kernel}_kernel (MC points)", "bound": "hbm", "achieved": alg_bytes / (k1_ms * 1e-3) / 1e9,
        "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
        "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes, "alg_bytes_per_eval": alg_bytes / evals_k1,
        "kernel_ms": k1_ms, "evals_per_s_kernel": evals_k1 / (k1_ms * 1e-3),
        "note": "The path is NOT HBM-bound (SURVEY 8d): compulsory traffic is ~2e-4 B per tree-eval, so the HBM "
                "fraction is tiny by construction; the binding resources are issue slots and shared-memory wavefronts "
                "(onchip_* fields; model in DESIGN.md section 3).",
        "equiv_stream_gbs": 32.0 * evals_k1 / (k1_ms * 1e-3) / 1e9,
        "hbm_fraction_equiv_stream": 32.0 * evals_k1 / (k1_ms * 1e-3) / 1e9 / hbm_peak,
        "smem_bound_evals_per_s": lsu_bound, "issue_bound_evals_per_s": issue_bound,
        "onchip_bound_evals_per_s": onchip_bound, "onchip_bound_fraction": evals_k1 / (k1_ms * 1e-3) / onchip_bound,
        "smem_bound_fraction": evals_k1 / (k1_ms * 1e-3) / lsu_bound,
        "sm_clock_mhz": sm_mhz, "trees_per_group": trees_per_group, "table_bits": nb, "fast_kernel": f0.fast_kernel,
    }
    # DRAM traffic per launch from the committed ncu --set full captures (profiles/ncu_traffic.json), same workload
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        traffic = {}
    if not args.points and not args.draws:
        roofline["traffic"] = traffic.get(f"eval_{f0.fast_kernel}_kernel")
    k0 = float(np.median(k0_ms))
    pm = f0.fast_kernel == "gather"                       # one pass: fp64 matrix -> 16 B packed record per point
    k0_name = "bin_pack_points_kernel" if pm else "bin_points_kernel"
    k0_bytes = (8 * N * d + 16 * N) if pm else 9 * N * d     # fp64 in; packed records (gather) or uint8 rows (table) out
    roofline_k0 = {"kernel": k0_name, "bound": "hbm", "achieved": k0_bytes / (k0 * 1e-3) / 1e9,
                   "peak": hbm_peak, "unit": "GB/s", "frac": k0_bytes / (k0 * 1e-3) / 1e9 / hbm_peak,
                   "traffic": traffic.get(k0_name) if not args.points else None, "kernel_ms": k0,
                   "alg_bytes_per_launch": k0_bytes, "peak_source": peak_src,
                   "note": "timed with events around the whole points_from_device call (allocation + launch gap included)"}

    # ---- CPU baseline beside it (oracle port, bounded sample, this box's host cores) ----
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import oracle_c as oc
        nthreads = host_cores(oc)
        rate, rows, secs = cpu_sample_run(flats[0], X_host.numpy().T, nthreads, args.cpu_budget, oc)
        rate1, rows1, _ = cpu_sample_run(flats[0], X_host.numpy().T, 1, min(args.cpu_budget, 5.0), oc)
        cpu = {"value": rate, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": f"{rows} of {N} MC points x all {S} draws x {m} trees ({secs:.1f} s), OpenMP over draws",
               "single_thread_value": rate1}

    # ---- the step's results against the CPU oracle's values for the same seeded inputs at FULL size ----
    # (tests/golden/bench_cfg2_expected.json, made by tools/bench_expected_cpu.py with oracle/oracle_c.c; the exact
    #  integrals and the chosen candidate do not depend on the number of ranks, the Monte Carlo mean does)
    check = {"integral_mean": float(d_int_ms[0].item()), "integral_sd": float(d_int_ms[1].item()),
             "mc_mean": float(results["mean_var"][0].item()), "argmax_candidate": int(torch.argmax(results["var"]).item())}
    golden_verdict = None
    if not args.points and not args.draws:
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_cfg2_expected.json")))["forests"]
            want = gold[str(2000 + (args.steps - 1) % len(forests))]
            keys = ["integral_mean", "integral_sd", "argmax_candidate"] + (["mc_mean"] if world == 1 else [])
            worst = max(abs(check[k] - want[k]) / max(abs(want[k]), 1e-300) for k in keys)
            golden_verdict = {"keys": keys, "max_rel_diff": worst, "agree_1e-9": bool(worst <= 1e-9)}
        except Exception as e:                      # a missing golden file must not cost the bench line
            golden_verdict = {"error": repr(e)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_json(c, {"mc_points_total": n_total,
                                                                       "parallelism": f"points x{world} (MC), draws x{world} (candidates)",
                                                                       "host_threads_per_rank": host_threads,
                                                                       # 0 = every tree evaluated point by point (the
                                                                       # headline); 1 / 2 = opt-in counting mode
                                                                       "count_splits": int(os.environ.get("BARTINT_COUNT_SPLITS", "0") or 0)}),
        "s_per_sequential_design_step": total_ms / args.steps / 1e3,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_k0": roofline_k0,
        "cpu_baseline": cpu,
        "results_check": dict(check, vs_cpu_oracle_full_size=golden_verdict),
    }
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: str) -> None:
    """Print the ONE result line on the real stdout (fd 1 is pointed at stderr for the rest of the run so that
    library chatter such as NCCL's version banner cannot pollute it)."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=0, help="override MC points per GPU (testing)")
    ap.add_argument("--draws", type=int, default=0, help="override posterior draws (testing)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-separate", action="store_true",
                    help="time the end-to-end step as separate ABI calls instead of the fused bartint_design_step call")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
