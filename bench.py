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


def reference_rows(c):
    """Rows of the MC matrix the CPU arm evaluates per step (a bounded sample of the workload; the rate is reported)."""
    return min(c["N"], 20000 * max(1, host_core_count() // 4))


def host_core_count():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def config_json(c, world):
    """The workload, identical in both arms (implementation details go to `impl_config`)."""
    return {"workload": c["name"], "d": c["d"], "trees": c["m"], "draws": c["S"], "mc_points_per_gpu": c["N"],
            "mc_points_total": c["N"] * world, "candidates": c["C"], "measure": c["measure"], "cutpoints_per_variable": 100,
            "l2": "GPU arm: flushed between steps (512 MiB memset outside the per-step event pairs)",
            "reference_arm_sample": f"the CPU arm (--impl reference) times {reference_rows(c)} of the {c['N']} MC points + all "
                                    f"{c['C']} candidates x all {c['S']} draws per step and reports the rate (linear in the points)"}


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
    rows = reference_rows(c)
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
        "config": config_json(c, args.gpus),
        "impl_config": {"reference_note": "R/dbarts cannot run here (no R); this is the C restatement of the dbarts predict "
                                          "loop + sampleIntegrals arithmetic (oracle/oracle_c.c), OpenMP over draws"},
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
def _event_pair():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def strong_line(name, forest, make_points, n_total, rank, world, dev, steps, warmup, flush):
    """A FIXED problem (n_total points x the forest) split over the ranks by points: every rank evaluates its slice with
    the bit-sliced kernel, the per-draw partial sums are all-gathered and added in rank order.  Device-timed, max over
    ranks.  make_points(lo, hi) -> Points of the rank's slice (untimed: the inputs are resident when the timing starts)."""
    import torch
    import torch.distributed as dist
    from bartint_b200 import dist as bdist
    lo, hi = bdist.shard_range(n_total, world, rank)
    pts = make_points(lo, hi)
    st = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    times, kms = [], []
    for i in range(warmup + steps):
        flush.zero_()
        if world > 1:
            dist.barrier()
        a, b = _event_pair()
        a.record()
        dm, mv = bdist.point_sharded_draw_means(forest, pts, stream=st, n_total=n_total)
        b.record(); b.synchronize()
        if i >= warmup:
            times.append(a.elapsed_time(b)); kms.append(forest.last_eval_profile()["kernel_ms"])
    ms = _max_over_ranks(float(np.mean(times)), dev, world)
    mean = float(mv[0].item())
    pts.close()
    evals = forest.S * forest.m * n_total
    return {"workload": name, "scaling": "strong", "points_total": n_total, "points_per_gpu": hi - lo, "draws": forest.S,
            "trees": forest.m, "ms_per_step": ms, "value": evals / (ms * 1e-3), "unit": UNIT, "kernel_ms_rank0": float(np.mean(kms)),
            "kernel": forest.last_eval_profile()["kernel"], "mc_mean": mean,
            "collective": "none" if world == 1 else "all_gather_into_tensor (NCCL) of the per-draw partial sums + rank-ordered sum"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import bartint_b200 as bi
    from bartint_b200 import Forest, synth, _lib
    from bartint_b200 import dist as bdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")      # host-side waits that must not put a spinning kernel on a GPU
    lib = _lib.load()
    # torchrun pins OMP_NUM_THREADS=1 per rank, which would make the exporter (host C++, OpenMP) serial: give every
    # rank its share of the host cores instead
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    n_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    host_threads = int(lib.bartint_set_host_threads(max(1, n_cores // max(local_world, 1))))
    c = workload(args)
    S, m, d, N, C = c["S"], c["m"], c["d"], c["N"], c["C"]
    n_total = N * world
    full_size = not args.points and not args.draws

    # ---- inputs (synthetic, seeded): posteriors of two consecutive design steps, MC points, candidates ----
    x_train, y_train = synth.make_design(CFG)
    flats = [synth.prior_forest(2000 + it, S, m, x_train, y_train) for it in range(2)]
    forests = [Forest.from_soa(ff) for ff in flats]
    assert all(f.table_kernel_ok and f.bitslice_ok for f in forests), "the fast kernels must apply to the benchmark forest"
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

    def step(i, no_bitslice=False):
        f = forests[i % len(forests)]
        # the large MC evaluation is enqueued first; the small chains (exact integral; candidates) follow on a side
        # stream so neither their kernels nor their host-side launch work sit in front of it
        side.wait_stream(main)
        pm = f.points_from_device(dX.data_ptr(), N, stream=st)
        dm, mv = bdist.point_sharded_draw_means(f, pm, stream=st, n_total=n_total, no_bitslice=no_bitslice)     # bartMean.R:74-82
        pm.close()
        with torch.cuda.stream(side):
            f.exact_integrals_device(d_int.data_ptr(), "uniform", d_rescaled=d_int_res.data_ptr(),
                                     d_mean_sd=d_int_ms.data_ptr(), stream=side.cuda_stream)    # sampleIntegrals + :260-262
            pc = f.points_from_device(dXc.data_ptr(), C, stream=side.cuda_stream)
            var = bdist.draw_sharded_candidate_var(f, pc, None)                                  # :312-314
            pc.close()
        main.wait_stream(side)
        results.update(var=var, draw_mean=dm, mean_var=mv, forest=f)

    def timed_steps(n_steps, n_warm, **kw):
        for i in range(n_warm):
            step(i, **kw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [_event_pair() for _ in range(n_steps)]
        kms = []
        torch.cuda.synchronize()
        for i in range(n_steps):
            flush.zero_()                                   # evict the previous step's data from the 126 MB L2
            ev[i][0].record()
            step(i, **kw)
            ev[i][1].record()
            ev[i][1].synchronize()
            kms.append(results["forest"].last_eval_profile()["kernel_ms"])    # the MC-point evaluation kernel
        torch.cuda.synchronize()
        return _max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev, world), kms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # samples from warm-up to the end of the timed region
    launches0 = lib.bartint_kernel_launch_count()
    total_ms, kernel_ms = timed_steps(args.steps, args.warmup)
    launches = (lib.bartint_kernel_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    check = {"integral_mean": float(d_int_ms[0].item()), "integral_sd": float(d_int_ms[1].item()),
             "mc_mean": float(results["mean_var"][0].item()), "argmax_candidate": int(torch.argmax(results["var"]).item())}
    mc_kernel = results["forest"].last_eval_profile()["kernel"]
    evals_per_step = S * m * (n_total + C)
    value = evals_per_step * args.steps / (total_ms * 1e-3)
    # the same step with the per-draw sums from the per-point (gather / table) kernel instead of the bit-sliced one
    pp_total_ms, pp_kernel_ms = timed_steps(max(2, min(args.steps, 5)), 1, no_bitslice=True)
    pp_steps = max(2, min(args.steps, 5))
    pp_kernel = results["forest"].last_eval_profile()["kernel"]

    # ---- K0 (binning) on its own: the HBM-bound kernel of the path; at the bench size and at 1.25e7 points ----
    def time_k0(dptr, n_rows, reps=5):
        out = []
        for _ in range(reps):
            flush.zero_()
            a, b = _event_pair()
            a.record()
            p = forests[0].points_from_device(dptr, n_rows, stream=st)
            b.record(); b.synchronize()
            out.append(a.elapsed_time(b))
            p.close()
        return float(np.median(out))
    k0 = time_k0(dX.data_ptr(), N)
    k0_large = None
    if full_size and rank == 0:
        n_big = 12_500_000                                   # cfg 4's per-GPU shard at 8 GPUs
        big = torch.rand((d, n_big), dtype=torch.float64, device=dev)
        k0_large = (n_big, time_k0(big.data_ptr(), n_big, 4))
        del big

    # ---- strong scaling (fixed problems): cfg 2's 1e6 points, and BASELINE cfg 4 = 1e8 gaussian-measure points ----
    strong = {}
    if not args.no_strong:
        f2 = forests[0]
        n2 = N                                                    # the one-GPU problem, split over the ranks
        Xs = synth.make_points(CFG, n2, seed=12345) if world > 1 else None

        def pts2(lo, hi):
            if world == 1:
                return f2.points_from_device(dX.data_ptr(), N, stream=st)
            shard = torch.from_numpy(np.asfortranarray(Xs[lo:hi]).T.copy()).to(dev)
            p = f2.points_from_device(shard.data_ptr(), hi - lo, stream=st)
            torch.cuda.synchronize()
            return p
        strong["cfg2_fixed_1e6_points"] = strong_line(c["name"], f2, pts2, n2, rank, world, dev, max(3, min(args.steps, 10)), 2, flush)
        if full_size:
            c4 = synth.CONFIGS[4]
            x4, y4 = synth.make_design(4)
            f4 = Forest.from_soa(synth.prior_forest(4000, c4["S"], c4["m"], x4, y4))
            n4 = args.cfg4_points
            pts4 = lambda lo, hi: f4.generate_points("gaussian", hi - lo, seed=4, first_index=lo, stream=st)   # noqa: E731
            line4 = strong_line(c4["name"], f4, pts4, n4, rank, world, dev, 3, 1, flush)
            line4["points"] = "generated on the device (Philox4x32-10, one global stream, each rank its slice)"
            if world > 1 and rank == 0:
                # the same fixed problem on ONE GPU of this box, so the line carries its own strong-scaling efficiency
                pts = f4.generate_points("gaussian", n4, seed=4, first_index=0, stream=st)
                sums = torch.empty(f4.S, dtype=torch.float64, device=dev)
                t1 = []
                for i in range(3):
                    flush.zero_()
                    a, b = _event_pair()
                    a.record()
                    f4.eval_device(pts, d_draw_sum=sums.data_ptr(), stream=st)
                    b.record(); b.synchronize()
                    if i:
                        t1.append(a.elapsed_time(b))
                pts.close()
                line4["one_gpu_same_box_ms"] = float(np.mean(t1))
                line4["strong_scaling_efficiency"] = float(np.mean(t1)) / (world * line4["ms_per_step"])
            if world > 1:
                dist.barrier()
            # sustained clocks: the 1-GPU cfg-4 evaluation back to back for >= 2 s with its own clock record
            if world == 1:
                pts = f4.generate_points("gaussian", n4, seed=4, first_index=0, stream=st)
                sums = torch.empty(f4.S, dtype=torch.float64, device=dev)
                f4.eval_device(pts, d_draw_sum=sums.data_ptr(), stream=st)
                torch.cuda.synchronize()
                smp = ClockSampler(local_rank)
                smp.start()
                a, b = _event_pair()
                reps = 0
                t0 = time.perf_counter()
                a.record()
                while time.perf_counter() - t0 < 2.5:
                    f4.eval_device(pts, d_draw_sum=sums.data_ptr(), stream=st)
                    torch.cuda.synchronize()
                    reps += 1
                b.record(); b.synchronize()
                sus_ms = a.elapsed_time(b) / reps
                line4["sustained"] = {"seconds": a.elapsed_time(b) / 1e3, "launches": reps, "ms_per_launch": sus_ms,
                                      "value": f4.S * f4.m * n4 / (sus_ms * 1e-3), "unit": UNIT, "clocks": smp.stop()}
                pts.close()
            f4.close()
            strong["cfg4_1e8_gaussian_points"] = line4

    # ---- end to end through the reference-facing call: dbarts wire format + host matrices in, host results out.
    #      ONE process drives all GPUs (the reference's caller is a single R process): rank 0 opens the library's device
    #      group over the `world` GPUs and calls bartint_design_step_dbarts098 with the whole N x world matrix staged on
    #      the group; the other ranks wait on the host. ----
    e2e = None
    if not args.no_e2e:
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            from bartint_b200 import StagedMatrix, WireStrings, design_step_dbarts
            lib.bartint_set_host_threads(n_cores)              # the other ranks are idle: the exporter gets the host
            states = []
            for ff in flats:
                strings, fits = ff.to_dbarts_state(x_train)
                states.append((WireStrings(strings), fits))    # R holds savedTrees as C strings already: marshal outside the timing
            cut_lists = flats[0].cut_lists()
            if world > 1:
                Xall = torch.empty((d, n_total), dtype=torch.float64).pin_memory()
                for r in range(world):
                    Xall[:, r * N:(r + 1) * N] = torch.from_numpy(synth.make_points(CFG, N, seed=12345 + r).T)
            else:
                Xall = X_host
            Xn, Xcn = Xall.numpy().T, Xc_host.numpy().T        # N x d views, column-major (pinned)
            if world > 1:
                assert bi.init_group(world) == world      # (keeps world - 1 cores free of exporter threads for its workers)
            target = "group" if world > 1 else local_rank
            e2e_steps = max(1, min(args.steps, 20))
            times, worst = [], 0.0
            gold = None
            if full_size and world == 1:
                try:
                    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_cfg2_expected.json")))["forests"]
                except Exception:
                    gold = None
            for k in range(e2e_steps + 2):
                wire, fits = states[k % 2]                     # a fresh posterior every step
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                # the point matrices do not depend on the fitted model: their H2D copies are started first and overlap
                # the exporter's CPU work (bartint_stage_matrix); everything goes through the host-buffer C ABI
                sx, sxc = StagedMatrix(Xn, target), StagedMatrix(Xcn, local_rank)
                o = design_step_dbarts(wire, fits, x_train, cut_lists, flats[k % 2].y_min, flats[k % 2].y_max, mc=sx,
                                       candidates=sxc, weights=1.0, measure="uniform")
                dt = time.perf_counter() - t0
                sx.close(); sxc.close()
                if k >= 2:
                    times.append(dt)
                if gold is not None:                           # the e2e loop's own outputs against the CPU oracle's
                    w = gold[str(2000 + k % 2)]
                    got = {"integral_mean": o["integral_mean_sd"][0], "integral_sd": o["integral_mean_sd"][1],
                           "mc_mean": o["draw_mean_var"][0], "argmax_candidate": int(np.argmax(o["cand_var"]))}
                    worst = max(worst, max(abs(got[q] - w[q]) / max(abs(w[q]), 1e-300) for q in w))
            f0 = forests[0]
            h2d = Xn.nbytes + Xcn.nbytes + f0.n_nodes * 9 + f0.n_leaves * 8 + f0.n_groups * 8 * (f0.table_bits + 1) + S * m * 8
            d2h = S * 8 * 2 + C * 8 + 48
            t_mean = float(np.mean(times))
            e2e = {"value": S * m * (n_total + C) / t_mean, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "s_per_step": t_mean, "s_per_step_min": float(np.min(times)),
                   "steps": e2e_steps, "gpus_driven_by_one_process": world,
                   "call": "bartint_stage_matrix x2 + bartint_design_step_dbarts098 (C ABI, host buffers)" +
                           (f"; MC matrix staged on the device group of {world} (bartint_init)" if world > 1 else ""),
                   "exchange": bi.group_exchange() if world > 1 else "none (one device)",
                   "includes": "X / candidate H2D from pinned memory (started first, overlapping the exporter), dbarts-0.9-8 "
                               "string/fit exporter (a fresh posterior every step), forest H2D + table build + bit-sliced "
                               "rows, image replication to the other GPUs, binning, exact integrals, evaluation, the "
                               "exchange, D2H of integrals, candidate variances and per-draw means",
                   "vs_cpu_oracle_full_size": None if gold is None else {"max_rel_diff": worst, "agree_1e-9": bool(worst <= 1e-9)}}
            if world > 1:
                bi.shutdown_group()
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (MC-point evaluation by the bit-sliced kernel) ----
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
    f0 = forests[0]
    bs = f0.bitslice_info
    tile = f0.bitslice_tile
    n_tiles = (N + tile - 1) // tile
    # Binding resource: the shared-memory data pipe (1 wavefront = 128 B per clock per SM).  A node at depth k >= 1 ANDs
    # k + 1 predicate rows of tile/8 bytes each, read by one lane with tile/128 LDS.128: tile/1024 wavefronts per row
    # reference, no way around it in this layout; root nodes use the per-tile count table instead.  bs[13] counts the
    # row references of the whole forest, bs[14] the references the padded slice rows actually load.
    wf_floor = n_tiles * bs[13] * (tile / 1024.0)
    wf_padded = n_tiles * bs[14] * (tile / 1024.0)
    smem_ceiling = evals_k1 / (wf_floor / (148 * sm_mhz * 1e6))
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        traffic = {}
    achieved = evals_k1 / (k1_ms * 1e-3)
    roofline = {
        "kernel": f"eval_{mc_kernel}_kernel (MC points, per-draw sums)", "bound": "smem",
        "achieved": achieved, "peak": smem_ceiling, "unit": UNIT, "frac": achieved / smem_ceiling,
        "traffic": traffic.get("eval_bitslice_kernel") if full_size else None,
        "peak_source": "shared-memory data pipe: 1 wavefront / clk / SM x 148 SMs x the SM clock sampled during the run; "
                       "wavefronts per launch from the forest's own node counts (DESIGN.md section 3)",
        "kernel_ms": k1_ms, "sm_clock_mhz": sm_mhz, "tile_points": tile, "tiles": n_tiles,
        "smem_wavefronts_floor": wf_floor, "smem_wavefronts_loaded_with_slice_padding": wf_padded,
        "frac_of_padded_ceiling": achieved / (evals_k1 / (wf_padded / (148 * sm_mhz * 1e6))) if wf_padded else None,
        "nodes_by_depth_0_1_2_3_4plus": bs[8:13],
        "hbm": {"bound": "hbm", "alg_bytes_per_launch": alg_bytes, "alg_bytes_per_eval": alg_bytes / evals_k1,
                "achieved": alg_bytes / (k1_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                "note": "not the binding resource (SURVEY 8d): ~2e-4 compulsory bytes per tree-eval"},
        "equiv_stream_gbs": 32.0 * achieved / 1e9, "hbm_fraction_equiv_stream": 32.0 * achieved / 1e9 / hbm_peak,
        "per_point_kernel": {"kernel": f"eval_{pp_kernel}_kernel (per-draw sums through the per-point kernel, BARTINT_NO_BITSLICE)",
                             "kernel_ms": float(np.mean(pp_kernel_ms)), "ms_per_step": pp_total_ms / pp_steps,
                             "evals_per_s_kernel": evals_k1 / (float(np.mean(pp_kernel_ms)) * 1e-3)},
    }
    pm = f0.fast_kernel == "gather"                       # one pass: fp64 matrix -> 16 B packed record per point
    k0_name = "bin_pack_points_kernel" if pm else "bin_points_kernel"

    def k0_line(rows, ms):
        nbytes = (8 * rows * d + 16 * rows) if pm else 9 * rows * d      # fp64 in; packed records or uint8 rows out
        return {"kernel": k0_name, "bound": "hbm", "points": rows, "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "kernel_ms": ms, "alg_bytes_per_launch": nbytes,
                "peak_source": peak_src,
                "note": "timed with events around the whole points_from_device call (allocation + launch gap included)"}
    roofline_k0 = k0_line(N, k0)
    roofline_k0["traffic"] = traffic.get(k0_name) if full_size else None
    if k0_large:
        roofline_k0["at_cfg4_shard_size"] = k0_line(*k0_large)

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
    golden_verdict = None
    if full_size:
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
        "dtype": "f64", "data": "synthetic",
        "config": config_json(c, world),
        "impl_config": {"parallelism": f"points x{world} (MC), draws x{world} (candidates)", "host_threads_per_rank": host_threads,
                        "mc_kernel": mc_kernel, "candidate_kernel": f0.fast_kernel},
        "s_per_sequential_design_step": total_ms / args.steps / 1e3,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_k0": roofline_k0,
        "cpu_baseline": cpu, "strong": strong,
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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling lines (fixed cfg 2 / cfg 4 problems)")
    ap.add_argument("--cfg4-points", type=int, default=100_000_000, help="points of the cfg-4 strong-scaling problem")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
