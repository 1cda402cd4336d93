#!/usr/bin/env python
"""bench.py -- CG iterations/s of the CORA staircase inner loop on the 100k-pose SE(3) RA-SLAM
problem (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = one bounded slice of the truncated-Newton solve: `--outer` trust-region iterations
(each one STPCG solve of at most 80 CG iterations + retraction + model/gradient evaluation)
continuing from the iterate the previous step left.  value = CG iterations of all ranks /
device time of the K timed steps (max over ranks), iterate resident in HBM.  e2e = the same
through cora_b200_tnt() with HOST (pinned) buffers: H2D of the iterate, the solve slice, D2H of
the result, every step.  Every rank solves its own random restart of the same problem (weak
scaling; no data-path collective -- the winning restart is gathered once at the end).

--impl reference times the CPU restatement of the reference (oracle/) on the host cores on a
bounded sample of the same workload; /root/reference cannot be built (no Eigen/SuiteSparse).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n=100_000, l=10, m=20_000, d=3, rank=5, seed=42)
METRIC = "cg_iterations_per_second"
UNIT = "CG it/s"


def algorithmic_bytes(nnz, N, r):
    """CSR-fp64-int32-equivalent traffic, independent of the storage format."""
    V = 8 * N * r
    q = 12 * nnz + 4 * (N + 1)
    return dict(spmm=q + 2 * V, hessvec=q + 4 * V, cg_iter=q + 15 * V + 8 * N)


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region (B200_PROFILING.md): NVML in a
    thread every 20 ms (nvidia-smi -lms as the fallback when pynvml is missing)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop_flag = index, [], None, None, False
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.dev, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                self.rows.append((float(sm), float(self.max_sm), int(mask)))
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                mask = 0
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        mask |= self.REASONS[name]
                self.rows.append((float(r[1]), float(r[2]), mask))
            except (ValueError, IndexError):
                continue

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no samples"]}
        sm = [r[0] for r in self.rows]
        reasons = sorted(k for k, bit in self.REASONS.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(r[1] for r in self.rows), "samples": len(sm),
                "reasons": reasons, "source": "nvml" if self.nvml else "nvidia-smi"}


def build_problem():
    from cora_b200 import capi, synthetic
    w = WORKLOAD
    arrays, gt = synthetic.make_arrays(w["n"], w["l"], w["m"], d=w["d"], seed=w["seed"])
    Q = capi.assemble(w["d"], w["n"], w["l"], arrays)
    m = len(arrays["rg_w"])
    return arrays, gt, Q, m


def initial_guess(arrays, gt, seed, kind):
    """warm: ground truth perturbed by Exp(N(0,0.05^2)) / N(0,0.5^2 m) -- the regime in which STPCG runs its
    full iteration budget; odom: the cold start of the reference's paper experiments
    (examples/paper_experiments.cpp:426-534); `seed` selects the restart (perturbation + SO(r) factor)."""
    from cora_b200 import synthetic
    w = WORKLOAD
    if kind == "odom":
        return synthetic.odometry_initialization(w["d"], w["n"], w["l"], arrays, w["rank"], seed=seed)
    return synthetic.perturbed_ground_truth(w["d"], w["n"], w["l"], arrays, gt, w["rank"], seed=seed)


def config_dict(args, N, nnz, m, impl="ours"):
    w = WORKLOAD
    cfg = {"workload": "synthetic 100k-pose SE(3) + 20k range factors (BASELINE configs[2]), rank %d" % w["rank"],
           "n_poses": w["n"], "n_landmarks": w["l"], "n_ranges": int(m), "N": int(N), "nnz": int(nnz),
           "rank": w["rank"], "preconditioner": "Jacobi",
           "init": {"warm": "ground truth perturbed (rotations 0.05 rad, positions 0.5 m), random SO(r) factor",
                    "odom": "odometry chain + random landmarks + random SO(r) factor "
                            "(paper_experiments.cpp:426-534)"}[args.init],
           "restarts": "one restart per GPU (seed = rank)"}
    if impl == "ours":
        cfg.update({"outer_iterations_per_step": args.outer, "untimed_startup_outer_iterations": args.pre_outer,
                    "max_TPCG_iterations": 80,
                    "l2": "working set of one CG iteration (Q 41 MB + 8 vectors x 16.8 MB = 175 MB) exceeds the 126 MB "
                          "L2; no explicit flush"})
    else:  # the CPU arm runs a BOUNDED SAMPLE of a step: what it runs is what it prints
        cfg.update({"outer_iterations_per_step": 1, "untimed_startup_outer_iterations": args.ref_pre,
                    "max_TPCG_iterations": args.ref_cg,
                    "sample_of": "one of the %d trust-region iterations of a GPU step, same problem, same initial "
                                 "guess, same STPCG budget" % args.outer})
    return cfg


def cpu_tnt_sample(arrays, gt, Q, m, args, steps, warmup, threads, min_seconds=0.0, reg_lambda=None):
    """The CPU restatement of the reference (oracle/cpu_ref.cpp) on a bounded sample of the workload:
    the same problem and initial guess, `ref_pre` untimed outer iterations, then steps of one TNT outer
    iteration with at most `ref_cg` CG iterations.  Returns (CG it/s, CG iterations, seconds, threads)."""
    from cora_b200 import capi
    from oracle import cpu_ref
    w = WORKLOAD
    if reg_lambda is None:
        R = cpu_ref.CpuRef(w["d"], w["n"], m, w["n"] + w["l"], Q, preconditioner=1, threads=threads)
    else:   # the reference's default preconditioner (RegularizedCholesky, same lambda as the GPU side)
        R = cpu_ref.CpuRef(w["d"], w["n"], m, w["n"] + w["l"], Q, preconditioner=3, reg_lambda=reg_lambda, threads=threads)
    x = R.project_to_manifold(initial_guess(arrays, gt, 0, args.init))
    pre = R.tnt(x, capi.default_tnt_params(max_iterations=args.ref_pre, max_TPCG_iterations=args.ref_cg,
                                           max_computation_time=0.0))
    x, delta = pre.x, pre.trust_region_radius[-1]
    its, T = 0, 0.0
    s = -1
    while True:
        s += 1
        if s >= warmup + steps and (min_seconds <= 0 or T >= min_seconds or s >= warmup + 40):
            break
        prm = capi.default_tnt_params(max_iterations=1, max_TPCG_iterations=args.ref_cg, Delta0=delta,
                                      max_computation_time=0.0)
        t0 = time.perf_counter()
        res = R.tnt(x, prm)
        dt = time.perf_counter() - t0
        x, delta = res.x, res.trust_region_radius[-1]
        if s >= warmup:
            its += int(sum(res.inner_iterations))
            T += dt
    return its / max(T, 1e-12), its, T, R.threads


# ------------------------------------------------------------------ reference arm ---
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arrays, gt, Q, m = build_problem()
    w = WORKLOAD
    N = w["d"] * w["n"] + m + w["n"] + w["l"]
    threads = os.cpu_count() or 1
    val, its, T, used = cpu_tnt_sample(arrays, gt, Q, m, args, args.steps, args.warmup, threads)
    sample = ("%d steps x (1 TNT outer iteration, <= %d CG iterations) of the same problem after %d untimed outer "
              "iterations; C++ restatement of the reference CPU path (oracle/cpu_ref.cpp: CSR, column-major, "
              "per-column SpMM, reference operation counts), %d threads; %d CG iterations in %.1f s"
              % (args.steps, args.ref_cg, args.ref_pre, used, its, T))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, N, Q.nnz, m, impl="reference"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference itself cannot be built on this image (needs Eigen3 + SuiteSparse): this is the "
                    "oracle's C++ port, kind=port"}
    print(json.dumps(line))


def roofline_cfg5(torch, dist, world, local, stream, peak, n=1_000_000, reps=30):
    """BASELINE configs[4]: synthetic 1M-pose SE(3) graph, rank 5.  HBM-roofline numbers of the data-matrix product
    (Problem::dataMatrixProduct, src/CORA_problem.cpp:742-757) and of full CG iterations.  The path does not shard a
    single solve: with N GPUs every rank runs its own replica and the aggregate is reported."""
    from cora_b200 import capi, synthetic
    d, r = 3, 5
    l, m = max(10, n // 10000), n // 5
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=42)
    Q = capi.assemble(d, n, l, arrays)
    m = len(arrays["rg_w"])
    N = d * n + m + n + l
    h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI, device=local, stream=stream)
    x0 = h.project_to_manifold(synthetic.perturbed_ground_truth(d, n, l, arrays, gt, r, seed=0))
    h.set_iterate(x0)
    ab = algorithmic_bytes(Q.nnz, N, r)
    h.spmm_resident(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ms = h.spmm_resident(reps)          # CUDA events around one launch that runs `reps` products back to back
    pre = h.tnt_resident(capi.default_tnt_params(max_iterations=12, max_computation_time=0.0))  # trust-region start-up
    res = h.tnt_resident(capi.default_tnt_params(max_iterations=2, max_computation_time=0.0,
                                                 Delta0=pre.trust_region_radius[-1]))
    clocks = sampler.stop()
    t_spmm = ms * 1e-3 / reps
    cg, outer = int(sum(res.inner_iterations)), len(res.inner_iterations)
    V = 8 * N * r
    b_outer = 2 * (12 * Q.nnz + 4 * (N + 1)) + 19 * V + 8 * N
    t_cg = res.device_time
    vals = torch.tensor([t_spmm, t_cg, float(cg)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_spmm, t_cg, cg_all = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        cg_all = float(cg)
    mxp, _ = ({}, {}) if not hasattr(h, "phase_profile_ctas") else (h.phase_profile_ctas(), None)
    # N > 1: the same product with ONE problem sharded by rows over the ranks (cora_b200/rowpart.py):
    # slabs of poses + ghost poses + replicated landmarks, exchanges by the library's kernels over peer-mapped memory
    rowp = None
    if world > 1:
        try:
            rowp = row_partitioned_cfg5(torch, dist, world, local, h, d, n, l, arrays, r, Q, reps)
        except Exception as e:  # never lose the bench line over the extra record
            rowp = {"error": "%s: %s" % (type(e).__name__, e)}
    h.close()
    traffic = None
    try:   # DRAM bytes per product from the committed ncu capture of this kernel on this workload (8 products per launch)
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r02_spmm_1m_ncu_raw.csv"))))
        hdr, units, v = rows[0], rows[1], rows[2]
        get = lambda name: float(v[hdr.index(name)].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index(name)]]
        traffic = (get("dram__bytes_read.sum") + get("dram__bytes_write.sum")) / 8.0
    except Exception:
        pass
    gbs_spmm = ab["spmm"] / t_spmm / 1e9
    gbs_cg = (cg * ab["cg_iter"] + outer * b_outer) / res.device_time / 1e9
    return {"workload": "synthetic %d-pose SE(3) + %d ranges, %d landmarks, rank %d (BASELINE configs[4]); one replica "
                        "per GPU" % (n, m, l, r),
            "N": int(N), "nnz": int(Q.nnz), "replicas": world, "row_partitioned": rowp,
            "spmm": {"us": 1e6 * t_spmm, "reps": reps, "algorithmic_bytes": ab["spmm"], "achieved_gbs": gbs_spmm,
                     "frac": gbs_spmm / peak, "aggregate_gbs": world * gbs_spmm, "traffic": traffic,
                     "traffic_kind": "committed ncu --set full capture (profiles/r02_spmm_1m_ncu_raw.csv), not measured in this run",
                     "timing": "CUDA events around one cooperative launch of `reps` products; max over ranks"},
            "cg_iteration": {"us": 1e6 * t_cg / max(1, cg), "iterations_timed": cg, "outer_iterations": outer,
                             "algorithmic_bytes": ab["cg_iter"], "achieved_gbs": gbs_cg, "frac": gbs_cg / peak,
                             "cg_it_per_s_all_replicas": cg_all / t_cg,
                             "phases_us_slowest_cta": {k: v[0] for k, v in mxp.items()}},
            "peak_gbs": peak, "clocks": clocks}


def row_partitioned_cfg5(torch, dist, world, local, h_full, d, n, l, arrays, r, Q, reps):
    """One 1M-pose problem sharded by rows over the ranks: time per data-matrix product and error against the full
    product of the rank's replica (h_full)."""
    from cora_b200 import capi, rowpart
    rank = dist.get_rank()
    parts = [rowpart.LocalProblem(d, n, l, arrays, world, g) for g in range(world)]
    P = parts[rank]
    X = np.random.default_rng(1).standard_normal((P.N, r))
    h_full.set_iterate(np.asfortranarray(X))
    h_full.spmm_resident(1)
    Yfull = h_full.get_work_vector(1, r)
    Ql = capi.assemble(d, P.n_loc, l, P.arrays)
    with capi.Handle(d, P.n_loc, P.m_loc, P.n_loc + l, Ql, preconditioner=capi.PRECON_JACOBI, device=local) as hl:
        Xl = X[P.local_to_global].copy()
        ghost = ~P.owned
        ghost[P.landmark_rows] = False
        Xl[ghost] = np.nan                      # only the exchange can make the product right
        hl.set_iterate(np.asfortranarray(Xl))
        pp, _ = rowpart.peer_product(hl, parts, rank, r, dist)
        dist.barrier()
        pp.product(1)
        Yl = hl.get_work_vector(1, r)
        rows = np.concatenate([np.nonzero(P.owned)[0], P.landmark_rows])
        err = float(np.abs(Yl[rows] - Yfull[P.local_to_global[rows]]).max() / np.abs(Yfull).max())
        pp.product(5)
        dist.barrier()
        us = 1e3 * pp.product(reps) / reps
        pp.close()
    v = torch.tensor([us, err], dtype=torch.float64, device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return {"what": "ONE problem sharded by rows over the ranks (SURVEY 8f-4): slab of poses + ghost poses + replicated "
                    "landmarks per rank; per product a cross-GPU flag barrier + pull of the ghost rows over peer-mapped "
                    "memory, the persistent SpMM kernel on the rank's rows, barrier + rank-ordered sum of the partial "
                    "landmark rows (cora_b200/csrc/peer_product.cuh) -- no collective",
            "n_gpus": world, "us_per_product": float(v[0]), "max_rel_error_vs_full_product": float(v[1]),
            "rows_per_rank": int(P.N_loc), "reps": reps,
            "timing": "CUDA events around `reps` products enqueued back to back; max over ranks"}


def spmv_cfg2(local, stream, reps=1000):
    """BASELINE configs[1]: Single-Drone SE(3), rank 5 -- the data-matrix product in isolation, `reps` repetitions,
    GPU (one cooperative launch) against the CPU restatement of Eigen's row-major-sparse x column-major-dense product
    (one pass over Q per column; 1 thread as the reference runs it, and all host threads)."""
    from cora_b200 import capi
    from oracle import cpu_ref
    path = os.path.join(ROOT, "tests", "golden", "single_drone.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    d, n, l = int(g["d"]), int(g["n"]), int(g["l"])
    arrays = {k: g[k] for k in g.files if k not in ("d", "n", "l")}
    Q = capi.assemble(d, n, l, arrays)
    m = len(arrays["rg_w"])
    r = 5
    N = d * n + m + n + l
    X = np.asfortranarray(np.random.default_rng(0).standard_normal((N, r)))
    with capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI, device=local, stream=stream) as h:
        h.set_iterate(X)
        h.spmm_resident(10)
        gpu_us = 1e3 * h.spmm_resident(reps) / reps
        got = h.get_work_vector(1, r)
    out = {"workload": "Single-Drone SE(3) (examples/data/single_drone.pyfg), rank 5, N = %d, nnz = %d" % (N, Q.nnz),
           "reps": reps, "gpu_us_per_product": gpu_us}
    for threads in (1, os.cpu_count() or 1):
        R = cpu_ref.CpuRef(d, n, m, n + l, Q, preconditioner=1, threads=threads)
        R.data_matrix_product(X)
        t0 = time.perf_counter()
        for _ in range(reps):
            ref = R.data_matrix_product(X)
        us = 1e6 * (time.perf_counter() - t0) / reps
        out["cpu_us_per_product_%s" % ("1_thread" if threads == 1 else "all_threads")] = us
        out["cpu_threads_all"] = R.threads
        R.close()
    out["max_rel_diff_gpu_vs_cpu"] = float(np.abs(got - ref).max() / np.abs(ref).max())
    out["speedup_vs_1_thread"] = out["cpu_us_per_product_1_thread"] / gpu_us
    return out


# ------------------------------------------------------------------------ our arm ---
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cora_b200 import build as _build, capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        _build.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if world > 1:
        dist.barrier()
    arrays, gt, Q, m = build_problem()
    w = WORKLOAD
    d, n, l, r = w["d"], w["n"], w["l"], w["rank"]
    N = d * n + m + n + l
    stream = torch.cuda.current_stream().cuda_stream
    h = capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI, device=local, stream=stream)
    ab = algorithmic_bytes(Q.nnz, N, r)
    prm = capi.default_tnt_params(max_iterations=args.outer, max_computation_time=0.0)

    x0 = h.project_to_manifold(initial_guess(arrays, gt, rank, args.init))
    pin_in = torch.empty((r, N), dtype=torch.float64).pin_memory()   # column-major N x r
    pin_out = torch.empty((r, N), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        its = launches = outer = 0
        prof = {}
        for _ in range(nsteps):
            a, b, c, p = step_fn()
            its += a
            launches += b
            outer += c
            for k, (us, cnt) in (p or {}).items():
                u0, c0 = prof.get(k, (0.0, 0))
                prof[k] = (u0 + us, c0 + cnt)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) * 1e-3, its, launches, outer, prof

    # ---- untimed: leave the start-up phase (the trust region grows from Delta0 = 5; STPCG ends on
    # its boundary after 0-3 iterations until then) so that the timed steps are CG iterations ----
    h.set_iterate(x0)
    pre = h.tnt_resident(capi.default_tnt_params(max_iterations=args.pre_outer, max_computation_time=0.0))
    delta_warm = pre.trust_region_radius[-1]
    h.snapshot_iterate()
    pin_warm = torch.empty((r, N), dtype=torch.float64).pin_memory()
    h.get_iterate_ptr(r, pin_warm.data_ptr())

    # ---- value leg: iterate resident in HBM ----
    prm.Delta0 = delta_warm
    dev_time = [0.0]

    def step_resident():
        res = h.tnt_resident(prm)
        prm.Delta0 = res.trust_region_radius[-1]   # the solve continues: keep the trust region
        if res.status != "IterationLimit":        # converged inside the bench: start the slice again
            h.restore_iterate()
            prm.Delta0 = delta_warm
        dev_time[0] += res.device_time
        prof, _, _ = h.phase_profile()
        return int(sum(res.inner_iterations)), res.kernel_launches, len(res.inner_iterations), prof

    timed(step_resident, args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    dev_time[0] = 0.0
    T, its, launches, outer, prof = timed(step_resident, args.steps)
    clocks = sampler.stop()
    t_kernel = dev_time[0]   # CUDA events on the launching stream around the persistent kernel of every step
    _, grid, _ = h.phase_profile()
    mxp = h.phase_profile_ctas()   # every CTA's clock for the same (last timed) launch as `prof`

    # ---- e2e leg: host buffers in, host buffers out, every step ----
    lib = capi.load()
    import ctypes as C
    resC, keep = capi.Handle._alloc_result(prm.max_iterations + 2)
    pin_in.copy_(pin_warm)

    def step_e2e():
        code = lib.cora_b200_tnt(h._h, C.c_int(r), C.cast(pin_in.data_ptr(), capi._PD), C.byref(prm),
                                 C.cast(pin_out.data_ptr(), capi._PD), C.byref(resC))
        capi._check(code)
        if resC.status == capi.TNT_STATUS.index("IterationLimit"):
            pin_in.copy_(pin_out)   # the next slice continues from this result (host side)
            prm.Delta0 = keep["trust_region_radius"][resC.num_outer]
        else:
            pin_in.copy_(pin_warm)
            prm.Delta0 = delta_warm
        return int(resC.total_inner), int(resC.kernel_launches), int(resC.num_outer), None

    prm.Delta0 = delta_warm
    timed(step_e2e, args.warmup)
    Te, its_e, _, _, _ = timed(step_e2e, args.steps)

    # ---- solve-to-certificate (BASELINE metric, second half; configs[3] with N > 1): every rank runs the staircase
    # on ITS restart with the reference's default preconditioner, then the best certified solution is gathered over
    # NCCL.  Reported beside the CG throughput, not timed into it. ----
    solve_cert = gather = None
    if not args.no_solve:
        comm = None
        if world > 1:
            from cora_b200 import restarts
            comm = restarts.make_native_comm(dist, local)      # communicator created (and warmed) before the clock starts
        h.set_preconditioner(capi.PRECON_REG_CHOLESKY)
        reg_lambda = h.reg_lambda
        # one untimed certification of a random point: CUDA loads kernels lazily, and the first use of the
        # certification kernels (Cholesky test, shift search, Lanczos) cost 0.05-0.9 s of module loading inside the
        # first solve (scripts/solve_leg_repeat.py: 0.30 s in the first solve of a process, 0.03 s afterwards)
        tw = time.perf_counter()
        for rw in (x0.shape[1], d):  # the staircase rank and the rank of the refinement stage
            # (a full-rank random point: the sv-ratio shortcut does not fire, the eigen-search runs)
            h.certify_solution(h.project_to_manifold(np.random.default_rng(1).uniform(-1.0, 1.0, size=(x0.shape[0], rw))),
                               0.05, max(10, d + 2))
        torch.cuda.synchronize()
        t_cert_warm = time.perf_counter() - tw
        barrier()
        ts = time.perf_counter()
        out = h.solve(x0, max_rank=7, params=capi.default_tnt_params(max_computation_time=0.0))
        torch.cuda.synchronize()
        t_solve = time.perf_counter() - ts
        # the PSD test of S + eta I at the solution on its own (src/CORA_utils.cpp:33-57): the certificate proper,
        # also when the staircase itself stopped on the reference's sv-ratio short-circuit
        Yd = out["x"]
        eta = min(max(out["f"] * 5e-6, 1e-7), 1e-1)
        tp = time.perf_counter()
        cert = h.certify_solution(Yd, eta, max(10, d + 2))
        torch.cuda.synchronize()
        t_psd = time.perf_counter() - tp
        tall = torch.tensor([t_solve], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tall, op=dist.ReduceOp.MAX)
        solve_cert = {"seconds": t_solve, "seconds_all_ranks_max": float(tall[0]), "certified": bool(out["certified"]),
                      "refined_certified": bool(out["refined_certified"]),
                      "f": float(out["f"]), "lifted_f": float(out["lifted_f"]), "lifted_rank": int(out["lifted_rank"]),
                      "cg_iterations": int(out["total_cg_iterations"]), "preconditioner": "RegularizedCholesky",
                      "stages": [{"rank": s["rank"], "status": s["status"], "outer": s["outer"], "cg": s["cg"],
                                  "certified": s["certified"], "cert_branch": s["cert_branch"], "theta": s["theta"],
                                  "tnt_s": s["tnt_seconds"], "cert_s": s["cert_seconds"]} for s in out["stages"]],
                      "psd_test_of_refined_solution": {"certified": bool(cert.is_certified), "branch": h.last_cert_branch,
                                                       "seconds": t_psd, "eta": eta},
                      "note": "rank 5 staircase (max rank 7) + rounding + refinement through cora_b200_solve(), host "
                              "buffers in/out; restart seed = rank; the reference's own stopping rules (src/CORA.cpp:95-109); "
                              "one untimed certification call before the clock loads the certification kernels (CUDA lazy "
                              "module loading)",
                      "untimed_certification_warmup_s": t_cert_warm}
        if rank == 0:
            # the same staircase with relative_decrease_tolerance = stepsize_tolerance = 0: the rank-5 solve then runs on to
            # the optimum of the rank-5 relaxation, where the certificate is the Cholesky PSD test of S + eta I itself
            tt = time.perf_counter()
            outt = h.solve(x0, max_rank=7, params=capi.default_tnt_params(
                max_computation_time=0.0, relative_decrease_tolerance=0.0, stepsize_tolerance=0.0,
                gradient_tolerance=1e-3, preconditioned_gradient_tolerance=0.0))
            torch.cuda.synchronize()
            solve_cert["tight_stopping_rules"] = {
                "seconds": time.perf_counter() - tt, "lifted_f": float(outt["lifted_f"]), "f": float(outt["f"]),
                "certified": bool(outt["certified"]), "cg_iterations": int(outt["total_cg_iterations"]),
                "stages": [{"rank": s["rank"], "status": s["status"], "outer": s["outer"], "cg": s["cg"],
                            "certified": s["certified"], "cert_branch": s["cert_branch"], "tnt_s": s["tnt_seconds"],
                            "cert_s": s["cert_seconds"]} for s in outt["stages"]]}
        if rank == 0:
            # the certificate proper at full size: TNT at rank 5 with the tight rules, then the Cholesky test of
            # S + eta I on the resident iterate (PSD half of fast_verification), timed on its own
            h.set_iterate(x0)
            rt = h.tnt_resident(capi.default_tnt_params(
                max_iterations=250, max_computation_time=0.0, relative_decrease_tolerance=0.0, stepsize_tolerance=0.0,
                gradient_tolerance=1e-3, preconditioned_gradient_tolerance=0.0))
            eta5 = min(max(rt.f * 5e-6, 1e-7), 1e-1)
            times, verdicts = [], []
            for _ in range(3):
                torch.cuda.synchronize()
                tq = time.perf_counter()
                verdicts.append(h.psd_test(eta5, r=r))
                torch.cuda.synchronize()
                times.append(time.perf_counter() - tq)
            solve_cert["psd_test_rank5_tight"] = {
                "f": float(rt.f), "gradient_norm": float(rt.gradfx_norm), "status": rt.status, "eta": eta5,
                "is_psd": bool(verdicts[-1]), "ms_median": 1e3 * float(np.median(times)),
                "what": "Lambda blocks + certificate values + chain Cholesky of S + eta I (N = %d), all on the device; "
                        "wall clock of cora_b200_psd_test on the resident iterate" % N}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            # CPU side of the same solve (RegularizedCholesky restated in oracle/cpu_ref.cpp): a full CPU
            # solve-to-certificate of this problem takes ~20 minutes (scripts/cpu_solve_100k.py, profiles/), so the
            # default run times a BOUNDED SAMPLE of its truncated-Newton iterations and extrapolates -- labelled so
            ncpu = os.cpu_count() or 1
            vc, citc, cTc, usedc = cpu_tnt_sample(arrays, gt, Q, m, args, 1, 0, ncpu, min_seconds=10.0,
                                                  reg_lambda=reg_lambda)
            solve_cert["cpu"] = {
                "kind": "port", "cores": usedc, "cg_it_per_s_regularized_cholesky": vc,
                "sample": "%d CG iterations of the CPU restatement with RegularizedCholesky in %.1f s (same problem, same "
                          "start, same lambda)" % (citc, cTc),
                "seconds_extrapolated": int(out["total_cg_iterations"]) / max(vc, 1e-12),
                "extrapolation": "GPU staircase CG iterations / CPU CG rate; excludes the CPU's certification and rounding",
                "full_run_record": "profiles/r02_cpu_solve_100k.json (scripts/cpu_solve_100k.py)"}
        if world > 1:
            # the refined rank-d solution of this rank's restart is the resident iterate of its handle
            h.set_iterate(out["x"])
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tg = time.perf_counter()
            g0.record()
            good = bool(out["refined_certified"] or out["certified"])
            win, wf = h.gather_best_resident(comm, world, rank, float(out["f"]), good)
            g1.record()
            torch.cuda.synchronize()
            t_g = time.perf_counter() - tg
            xbest = h.get_iterate(d)
            same = bool(np.array_equal(xbest, out["x"])) if int(win) == rank else None
            chk = torch.tensor([float(np.abs(xbest).sum())], dtype=torch.float64, device="cuda")
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            tmax = torch.tensor([t_g], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            gather = {"winner_rank": int(win), "winner_f": float(wf), "ms": 1e3 * float(tmax[0]),
                      "ms_device_rank0": g0.elapsed_time(g1), "bytes_broadcast": int(8 * N * d),
                      "winner_iterate_bit_identical_on_winner": same,
                      "all_ranks_hold_the_same_iterate": bool(float(lo[0]) == float(hi[0])),
                      "transport": "ncclAllGather of {f, certified, rank} + ncclBroadcast of the winner's resident N x d "
                                   "iterate, device to device (library communicator, warmed at creation)"}
            comm.close()
        h.set_preconditioner(capi.PRECON_JACOBI)

    # ---- the two other BASELINE configurations the north star asks numbers for ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md) (of fallback)"
    h.close()
    cfg5 = None if args.no_cfg5 else roofline_cfg5(torch, dist, world, local, stream, peak)
    cfg2 = None if (args.no_cfg2 or rank != 0) else spmv_cfg2(local, stream)

    # ---- aggregate over ranks ----
    vals = torch.tensor([T, Te, float(its), float(its_e), float(launches)], dtype=torch.float64, device="cuda")
    its_rank0, outer_rank0 = its, outer
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        T, Te = float(mx[0]), float(mx[1])
        its, its_e, launches = float(sm[2]), float(sm[3]), float(sm[4])
    value = its / T
    e2e = its_e / Te

    if rank == 0:
        # algorithmic bytes of everything the kernel did in the timed steps: per CG iteration
        # B_cgiter, per outer iteration retract 4V + f/grad product Q+3V + model Hess-vec Q+4V +
        # preconditioned gradient 3V+8N + STPCG init 5V
        V = 8 * N * r
        q = 12 * Q.nnz + 4 * (N + 1)
        b_outer = 2 * q + 19 * V + 8 * N
        total_bytes = its_rank0 * ab["cg_iter"] + outer_rank0 * b_outer
        ach = total_bytes / t_kernel / 1e9
        # DRAM traffic: cannot be measured without a profiler attached.  What is printed is the traffic / algorithmic
        # ratio of the committed `ncu --set full` capture of the same kernel (one launch of a smaller slice of the same
        # workload), scaled to this run's algorithmic bytes per launch -- labelled as such.
        traffic, traffic_note = None, None
        for cap_file in ("r02_persistent_ncu_raw.csv", "r01f_persistent_ncu_raw.csv"):
            try:
                import csv
                rows = list(csv.reader(open(os.path.join(ROOT, "profiles", cap_file))))
                hdr, units, vals = rows[0], rows[1], rows[2]
                get = lambda name: float(vals[hdr.index(name)].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index(name)]]
                cap = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
                cap_alg = 160 * ab["cg_iter"] + 2 * b_outer      # the captured launch: 2 outer iterations, 160 CG iterations
                traffic = cap / cap_alg * total_bytes / max(1, args.steps)
                traffic_note = ("ncu --set full of one launch with 160 CG + 2 outer iterations (profiles/%s): %.2f GB DRAM "
                                "read+write vs %.2f GB algorithmic (ratio %.3f); scaled to this launch's algorithmic bytes"
                                % (cap_file, cap / 1e9, cap_alg / 1e9, cap / cap_alg))
                break
            except Exception:
                continue
        phases = {}
        for k, (us, cnt) in prof.items():
            if cnt:
                phases[k] = {"avg_us": us / cnt, "count": cnt}
        for k, (mx_us, md_us) in mxp.items():   # every CTA's clock: the phase lasts as long as its slowest CTA
            if k in phases:
                phases[k]["slowest_cta_avg_us"] = mx_us
                phases[k]["median_cta_avg_us"] = md_us
        if "hess" in phases:
            t_h = phases["hess"].get("slowest_cta_avg_us", phases["hess"]["avg_us"])
            phases["hess"]["algorithmic_bytes"] = ab["hessvec"]
            phases["hess"]["achieved_gbs"] = ab["hessvec"] / (t_h * 1e-6) / 1e9
            phases["hess"]["frac"] = phases["hess"]["achieved_gbs"] / peak
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, N, Q.nnz, m),
                "cg_iterations_timed": int(its), "us_per_cg_iteration": 1e6 * T * world / max(1.0, its),
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(8 * N * r),
                        "d2h_bytes_per_step": int(8 * N * r)},
                "gpu_launches": int(launches), "gather_best": gather, "solve_to_cert": solve_cert,
                "roofline_cfg5": cfg5, "spmv_cfg2": cfg2,
                "roofline": {"bound": "hbm",
                             "kernel": "k_tnt_persistent<3>: one cooperative launch per step runs the whole TNT slice "
                                       "(%d CG iterations + %d outer iterations per launch on average)"
                                       % (its_rank0 // max(1, args.steps), outer_rank0 // max(1, args.steps)),
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": total_bytes / max(1, args.steps),
                             "avg_launch_us": 1e6 * t_kernel / max(1, args.steps), "launches_timed": args.steps,
                             "traffic": traffic, "traffic_kind": "scaled from committed capture (not measured in this run)",
                             "traffic_source": traffic_note, "grid": grid,
                             "timing": "CUDA events on the launching stream around every launch of the timed steps",
                             "phases_in_kernel_globaltimer": phases,
                             "phases_note": "avg_us: CTA 0's clock; slowest_cta_avg_us / median_cta_avg_us: over all CTAs of "
                                            "the last timed launch"}}
        if not args.no_cpu_baseline and world == 1:
            ncpu = os.cpu_count() or 1
            v, cits, cT, used = cpu_tnt_sample(arrays, gt, Q, m, args, 2, 0, ncpu, min_seconds=10.0)
            v1, cits1, cT1, _ = cpu_tnt_sample(arrays, gt, Q, m, args, 1, 0, 1, min_seconds=6.0)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": used, "kind": "port",
                "value_1_thread": v1,
                "sample": "steps of (1 TNT outer iteration, <= %d CG iterations) for >= 10 s of the same problem and initial guess after "
                          "%d untimed outer iterations; C++ restatement of the reference CPU path (oracle/cpu_ref.cpp), "
                          "%d threads: %d CG iterations in %.1f s; the reference itself is single-threaded: 1 thread: %d CG "
                          "iterations in %.1f s" % (args.ref_cg, args.ref_pre, used, cits, cT, cits1, cT1)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--outer", type=int, default=10, help="TNT outer iterations per step")
    ap.add_argument("--init", default="warm", choices=["warm", "odom"])
    ap.add_argument("--pre-outer", type=int, default=12,
                    help="untimed TNT outer iterations before the first step (trust-region start-up)")
    ap.add_argument("--ref-cg", type=int, default=80, help="CG cap per outer iteration of the CPU sample (= the GPU arm's)")
    ap.add_argument("--ref-pre", type=int, default=8, help="untimed outer iterations before the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the solve-to-certificate leg")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the 1M-pose roofline leg (BASELINE configs[4])")
    ap.add_argument("--no-cfg2", action="store_true", help="skip the Single-Drone SpMV leg (BASELINE configs[1])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
