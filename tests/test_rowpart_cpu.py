"""Row-partitioned data-matrix product (cora_b200/rowpart.py) -- the partition and exchange logic on CPU.

The slab of every rank (owned poses + ghosts + replicated landmarks + owned ranges) is assembled with the oracle and
multiplied with SciPy; owned rows must reproduce the rows of the global product Q X (src/CORA_problem.cpp:742-757) and
the landmark rows must add up over the ranks.  The gloo world-size-2 test runs the real exchange code
(all_to_all_single of the ghost pose blocks, all_reduce of the landmark rows) with the SciPy product standing in for
the GPU kernel."""
import os
import socket

import numpy as np
import pytest

from cora_b200 import rowpart, synthetic
from oracle import cora_oracle as co


def _problem(n=240, l=3, m=150, d=3, loops=None):
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=5, loop_closures=loops)
    Q = co.assemble_Q(d, n, l, arrays)
    return d, n, l, arrays, Q


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("loops", [None, [(3, 200), (50, 120), (119, 121)]])
def test_slabs_reproduce_the_global_product(world, loops):
    d, n, l, arrays, Q = _problem(loops=loops)
    r = 4
    X = np.random.default_rng(0).standard_normal((Q.shape[0], r))
    Y = Q @ X
    parts = [rowpart.LocalProblem(d, n, l, arrays, world, g) for g in range(world)]
    covered = np.zeros(Q.shape[0], dtype=int)
    lm_sum = 0
    for P in parts:
        Ql = co.assemble_Q(d, P.n_loc, l, P.arrays)
        assert Ql.shape[0] == P.N_loc
        Yl = Ql @ X[P.local_to_global]
        own = P.owned
        np.testing.assert_allclose(Yl[own], Y[P.local_to_global[own]], rtol=0, atol=1e-9 * np.abs(Y).max())
        covered[P.local_to_global[own]] += 1
        lm_sum = lm_sum + Yl[P.landmark_rows]
        lm_glob = P.local_to_global[P.landmark_rows]
    assert np.all(covered[: d * n] == 1) and np.all(covered[d * n: d * n + parts[0].m] == 1)   # every pose / range row once
    np.testing.assert_allclose(lm_sum, Y[lm_glob], rtol=0, atol=1e-9 * np.abs(Y).max())
    # chain: one ghost per interior boundary side
    if loops is None and world > 1:
        assert [P.n_loc - P.n_own for P in parts] == [1] + [2] * (world - 2) + [1]


def test_cross_slab_pose_pose_range_is_rejected():
    d, n, l, arrays, _ = _problem()
    a = dict(arrays)
    a["rg_a"] = np.r_[arrays["rg_a"], 5]; a["rg_b"] = np.r_[arrays["rg_b"], 200]
    a["rg_r"] = np.r_[arrays["rg_r"], 1.0]; a["rg_w"] = np.r_[arrays["rg_w"], 1.0]
    with pytest.raises(NotImplementedError):
        rowpart.LocalProblem(d, n, l, a, 2, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, n, l, arrays, Q = _problem(loops=[(3, 200), (119, 121)])
        r = 3
        X = np.random.default_rng(0).standard_normal((Q.shape[0], r))
        Y = Q @ X
        parts = [rowpart.LocalProblem(d, n, l, arrays, world, g) for g in range(world)]
        P = parts[rank]
        send = rowpart.exchange_plan(parts)
        Ql = co.assemble_Q(d, P.n_loc, l, P.arrays)
        # the rank knows its owned rows and the (replicated) landmark rows; ghost rows start as garbage
        xl = X[P.local_to_global].copy()
        ghost = ~P.owned
        ghost[P.landmark_rows] = False
        xl[ghost] = np.nan
        perm = np.random.default_rng(7 + rank).permutation(P.N_loc)   # an arbitrary buffer row order, as on the device
        row_of = np.empty(P.N_loc, dtype=np.int64); row_of[perm] = np.arange(P.N_loc)
        xb = torch.from_numpy(xl[perm].copy()); yb = torch.zeros_like(xb)

        def product():
            yb.copy_(torch.from_numpy((Ql @ xb.numpy()[row_of])[perm]))

        op = rowpart.RowPartitionedProduct(P, send, xb, yb, row_of, product, dist)
        op()
        yl = yb.numpy()[row_of]
        rows = np.concatenate([np.nonzero(P.owned)[0], P.landmark_rows])
        err = float(np.abs(yl[rows] - Y[P.local_to_global[rows]]).max() / np.abs(Y).max())
        q.put((rank, err, bool(np.isfinite(xb.numpy()).all()), int(P.n_loc - P.n_own)))
    finally:
        dist.destroy_process_group()


def test_exchange_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, finite, nghost in res:
        assert err <= 1e-12, (rank, err)
        assert finite          # every ghost row was filled by its owner
        assert nghost >= 1
