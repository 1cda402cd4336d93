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


def config_dict(args, N, nnz, m):
    w = WORKLOAD
    return {"workload": "synthetic 100k-pose SE(3) + 20k range factors (BASELINE configs[2]), rank %d" % w["rank"],
            "n_poses": w["n"], "n_landmarks": w["l"], "n_ranges": int(m), "N": int(N), "nnz": int(nnz),
            "rank": w["rank"], "preconditioner": "Jacobi", "outer_iterations_per_step": args.outer,
            "untimed_startup_outer_iterations": args.pre_outer, "max_TPCG_iterations": 80,
            "init": {"warm": "ground truth perturbed (rotations 0.05 rad, positions 0.5 m), random SO(r) factor",
                     "odom": "odometry chain + random landmarks + random SO(r) factor "
                             "(paper_experiments.cpp:426-534)"}[args.init],
            "restarts": "one restart per GPU (seed = rank)",
            "l2": "working set of one CG iteration (Q 41 MB + 8 vectors x 16.8 MB = 175 MB) exceeds the 126 MB L2; "
                  "no explicit flush"}


def cpu_tnt_sample(arrays, gt, Q, m, args, steps, warmup, threads, min_seconds=0.0):
    """The CPU restatement of the reference (oracle/cpu_ref.cpp) on a bounded sample of the workload:
    the same problem and initial guess, `ref_pre` untimed outer iterations, then steps of one TNT outer
    iteration with at most `ref_cg` CG iterations.  Returns (CG it/s, CG iterations, seconds, threads)."""
    from cora_b200 import capi
    from oracle import cpu_ref
    w = WORKLOAD
    R = cpu_ref.CpuRef(w["d"], w["n"], m, w["n"] + w["l"], Q, preconditioner=1, threads=threads)
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
            "data": "synthetic", "config": config_dict(args, N, Q.nnz, m),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference itself cannot be built on this image (needs Eigen3 + SuiteSparse): this is the "
                    "oracle's C++ port, kind=port"}
    print(json.dumps(line))


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

    # ---- solve-to-certificate of the whole problem (BASELINE metric, second half): staircase from the same
    # start with the reference's default preconditioner; reported beside the CG throughput, not timed into it ----
    solve_cert = None
    if not args.no_solve and rank == 0:
        h.set_preconditioner(capi.PRECON_REG_CHOLESKY)
        torch.cuda.synchronize()
        ts = time.perf_counter()
        out = h.solve(x0, max_rank=7, params=capi.default_tnt_params(max_computation_time=0.0))
        torch.cuda.synchronize()
        solve_cert = {"seconds": time.perf_counter() - ts, "certified": bool(out["certified"]),
                      "f": float(out["f"]), "lifted_f": float(out["lifted_f"]), "lifted_rank": int(out["lifted_rank"]),
                      "cg_iterations": int(out["total_cg_iterations"]), "preconditioner": "RegularizedCholesky",
                      "stages": [{"rank": s["rank"], "status": s["status"], "outer": s["outer"], "cg": s["cg"],
                                  "certified": s["certified"], "tnt_s": s["tnt_seconds"], "cert_s": s["cert_seconds"]}
                                 for s in out["stages"]],
                      "note": "rank 5 -> 7 staircase + rounding + refinement through cora_b200_solve(), host buffers in/out"}
        h.set_preconditioner(capi.PRECON_JACOBI)

    # ---- the only exchange of the multi-GPU path: gather the best restart (outside the timed region) ----
    gather = None
    if world > 1:
        from cora_b200 import restarts
        xr = np.ascontiguousarray(pin_out.numpy().T)           # N x r, this rank's iterate after the e2e leg
        comm = restarts.make_native_comm(dist, local)
        barrier()
        tg = time.perf_counter()
        win, wf, _ = restarts.gather_best(dist, float(resC.f), False, xr, handle=h, comm=comm)
        torch.cuda.synchronize()
        gather = {"winner_rank": int(win), "winner_f": float(wf), "ms": 1e3 * (time.perf_counter() - tg),
                  "bytes_broadcast": int(8 * N * r), "transport": "ncclAllGather + ncclBroadcast (library communicator)"}
        comm.close()

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
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md) (of fallback)"
        # algorithmic bytes of everything the kernel did in the timed steps: per CG iteration
        # B_cgiter, per outer iteration retract 4V + f/grad product Q+3V + model Hess-vec Q+4V +
        # preconditioned gradient 3V+8N + STPCG init 5V
        V = 8 * N * r
        q = 12 * Q.nnz + 4 * (N + 1)
        b_outer = 2 * q + 19 * V + 8 * N
        total_bytes = its_rank0 * ab["cg_iter"] + outer_rank0 * b_outer
        ach = total_bytes / t_kernel / 1e9
        # DRAM traffic of the same kernel from the committed `ncu --set full` capture (one launch of a smaller slice):
        # the measured traffic / algorithmic ratio of that capture, scaled to this run's launch
        traffic, traffic_note = None, None
        try:
            import csv
            rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r01f_persistent_ncu_raw.csv"))))
            hdr, units, vals = rows[0], rows[1], rows[2]
            get = lambda name: float(vals[hdr.index(name)].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index(name)]]
            cap = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
            cap_alg = 160 * ab["cg_iter"] + 2 * b_outer          # the captured launch: 2 outer iterations, 160 CG iterations
            traffic = cap / cap_alg * total_bytes / max(1, args.steps)
            traffic_note = ("ncu --set full of one launch with 160 CG + 2 outer iterations (profiles/r01f_persistent_ncu_raw.csv): "
                            "%.2f GB DRAM read+write vs %.2f GB algorithmic (ratio %.3f); scaled to this launch's algorithmic bytes"
                            % (cap / 1e9, cap_alg / 1e9, cap / cap_alg))
        except Exception:
            pass
        phases = {}
        for k, (us, cnt) in prof.items():
            if cnt:
                phases[k] = {"avg_us": us / cnt, "count": cnt}
        if "hess" in phases:
            phases["hess"]["algorithmic_bytes"] = ab["hessvec"]
            phases["hess"]["achieved_gbs"] = ab["hessvec"] / (phases["hess"]["avg_us"] * 1e-6) / 1e9
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
                "roofline": {"bound": "hbm",
                             "kernel": "k_tnt_persistent<3>: one cooperative launch per step runs the whole TNT slice "
                                       "(%d CG iterations + %d outer iterations per launch on average)"
                                       % (its_rank0 // max(1, args.steps), outer_rank0 // max(1, args.steps)),
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": total_bytes / max(1, args.steps),
                             "avg_launch_us": 1e6 * t_kernel / max(1, args.steps), "launches_timed": args.steps,
                             "traffic": traffic, "traffic_source": traffic_note, "grid": grid,
                             "timing": "CUDA events on the launching stream around every launch of the timed steps",
                             "phases_in_kernel_globaltimer_cta0": phases}}
        if not args.no_cpu_baseline and world == 1:
            v, cits, cT, used = cpu_tnt_sample(arrays, gt, Q, m, args, 2, 0, os.cpu_count() or 1, min_seconds=10.0)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": used, "kind": "port",
                "sample": "steps of (1 TNT outer iteration, <= %d CG iterations) for >= 10 s of the same problem and initial guess after "
                          "%d untimed outer iterations; C++ restatement of the reference CPU path (oracle/cpu_ref.cpp), "
                          "%d threads; %d CG iterations in %.1f s" % (args.ref_cg, args.ref_pre, used, cits, cT)}
        print(json.dumps(line))
    h.close()
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
    ap.add_argument("--ref-cg", type=int, default=40, help="CG cap per outer iteration of the CPU sample")
    ap.add_argument("--ref-pre", type=int, default=8, help="untimed outer iterations before the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the solve-to-certificate leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
