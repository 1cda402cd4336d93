"""Row-partitioned data-matrix product across the GPUs of one node: correctness against the full product
on one GPU and time per product.  Launch: torchrun --nproc-per-node N scripts/rowpart_bench.py [n_poses] [reps] [out.json]
(N = 1 works too: no exchange)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cora_b200 import capi, rowpart, synthetic  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, r = 3, 5
    l, m = max(10, n // 10000), n // 5
    arrays, _ = synthetic.make_arrays(n, l, m, d=d, seed=42)
    m = len(arrays["rg_w"])
    parts = [rowpart.LocalProblem(d, n, l, arrays, world, g) for g in range(world)]
    P = parts[rank]
    send = rowpart.exchange_plan(parts)
    Ql = capi.assemble(d, P.n_loc, l, P.arrays)
    N = P.N
    X = np.random.default_rng(1).standard_normal((N, r))
    h = capi.Handle(d, P.n_loc, P.m_loc, P.n_loc + l, Ql, preconditioner=capi.PRECON_JACOBI, device=local)
    h.set_iterate(np.asfortranarray(X[P.local_to_global]))
    op, x, y, row_of = rowpart.device_product(h, P, send, r, dist if world > 1 else None)
    # ghost rows of the operand are poisoned: only the exchange can make the product right
    ghost = ~P.owned
    ghost[P.landmark_rows] = False
    if ghost.any():
        x[torch.as_tensor(row_of[np.nonzero(ghost)[0]], device=x.device)] = float("nan")
    op()
    torch.cuda.synchronize()
    # reference: the full product on this GPU
    Q = capi.assemble(d, n, l, arrays)
    with capi.Handle(d, n, m, n + l, Q, preconditioner=capi.PRECON_JACOBI, device=local) as hf:
        hf.set_iterate(np.asfortranarray(X))
        ms_full = hf.spmm_resident(reps) / reps
        Yfull = hf.get_work_vector(1, r)
    rows = np.concatenate([np.nonzero(P.owned)[0], P.landmark_rows])
    yl = y[torch.as_tensor(row_of[rows], device=y.device)].cpu().numpy()
    err = float(np.abs(yl - Yfull[P.local_to_global[rows]]).max() / np.abs(Yfull).max())
    # timing: whole product, and its parts
    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return (time.perf_counter() - t0) / reps * 1e6
    for _ in range(5):
        op()
    us_total = timed(op)
    us_halo = timed(op.halo)
    us_local = timed(op.product_fn)
    us_reduce = timed(op.reduce_landmarks)
    # the same product with the library's own exchange kernels over peer-mapped memory (no collective per product)
    pp, _ = rowpart.peer_product(h, parts, rank, r, dist if world > 1 else None)
    h.set_iterate(np.asfortranarray(X[P.local_to_global]))
    if ghost.any():
        x[torch.as_tensor(row_of[np.nonzero(ghost)[0]], device=x.device)] = float("nan")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    pp.product(1)
    yl2 = y[torch.as_tensor(row_of[rows], device=y.device)].cpu().numpy()
    err_peer = float(np.abs(yl2 - Yfull[P.local_to_global[rows]]).max() / np.abs(Yfull).max())
    pp.product(5)
    if world > 1:
        dist.barrier()
    us_peer = 1e3 * pp.product(reps) / reps
    pp.close()
    vals = torch.tensor([us_total, us_halo, us_local, us_reduce, err, us_peer, err_peer], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    if rank == 0:
        rec = {"workload": "synthetic %d-pose SE(3) + %d ranges, %d landmarks, rank %d" % (n, m, l, r), "n_gpus": world,
               "rows_per_rank": int(P.N_loc), "ghost_poses_this_rank": int(P.n_loc - P.n_own), "reps": reps,
               "us_per_product_row_partitioned": float(vals[0]), "us_halo_exchange": float(vals[1]),
               "us_local_product": float(vals[2]), "us_landmark_all_reduce": float(vals[3]),
               "us_per_product_one_gpu": 1e3 * ms_full, "max_rel_error_vs_full_product": float(vals[4]),
               "us_per_product_peer_memory_kernels": float(vals[5]), "max_rel_error_peer_memory_vs_full_product": float(vals[6]),
               "peer_memory_timing": "CUDA events around `reps` products enqueued back to back (barrier + halo pull | local rows | barrier + landmark sum), max over ranks",
               "timing": "wall clock around `reps` products between device synchronisations and barriers, max over ranks"}
        print(json.dumps(rec))
        if out_path:
            os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
            with open(out_path, "a") as fh:
                fh.write(json.dumps(rec) + "\n")
    assert float(vals[4]) <= 1e-12, float(vals[4])
    assert float(vals[6]) <= 1e-12, float(vals[6])
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
