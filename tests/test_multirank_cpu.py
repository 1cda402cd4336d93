"""N > 1 host logic on CPU (gloo, world_size 2): restart seeds, the winner rule of
cora_b200_gather_best and the gather/broadcast protocol.  No GPU, no compute calls."""
import os
import socket

import numpy as np
import pytest


def test_select_best_rule(lib):
    from cora_b200 import capi
    # certified beats uncertified even at higher cost; among certified the lowest cost wins
    assert capi.select_best([3.0, 1.0, 2.0], [1, 0, 1]) == 2
    # nobody certified: lowest cost overall
    assert capi.select_best([3.0, 1.0, 2.0], [0, 0, 0]) == 1
    # ties go to the lowest rank; NaN never wins
    assert capi.select_best([2.0, 2.0], [1, 1]) == 0
    assert capi.select_best([float("nan"), 5.0], [1, 1]) == 1
    with pytest.raises(capi.InvalidArgument):
        capi.select_best([], [])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from cora_b200 import restarts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seed = restarts.restart_seed(rank)
        rng = np.random.default_rng(seed)
        X = np.asfortranarray(rng.standard_normal((37, 4)))
        f = [5.0, 2.5][rank]          # rank 1 has the lower cost ...
        cert = [True, False][rank]    # ... but only rank 0 is certified -> rank 0 wins
        win, wf, Xw = restarts.gather_best(dist, f, cert, X)
        win2, wf2, Xw2 = restarts.gather_best(dist, f, False, X)   # nobody certified -> rank 1 wins
        q.put((rank, seed, win, wf, float(np.abs(Xw).sum()), win2, wf2, float(np.abs(Xw2).sum()),
               float(np.abs(X).sum())))
    finally:
        dist.destroy_process_group()


def test_gather_best_two_ranks_gloo(lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, w0, f0, x0, w0b, f0b, x0b, own0), (r1, s1, w1, f1, x1, w1b, f1b, x1b, own1) = res
    assert (s0, s1) == (0, 1)
    assert w0 == w1 == 0 and f0 == f1 == 5.0
    assert abs(x0 - own0) < 1e-12 and abs(x1 - own0) < 1e-12      # both hold rank 0's iterate
    assert w0b == w1b == 1 and f0b == f1b == 2.5
    assert abs(x0b - own1) < 1e-12 and abs(x1b - own1) < 1e-12    # both hold rank 1's iterate
