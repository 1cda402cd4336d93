"""Row-partitioned data-matrix product of ONE problem across GPUs.

The reference multiplies `data_matrix_ * Y` on one core (src/CORA_problem.cpp:742-757).  Here rank g owns a
contiguous range of poses -- in the pose-major device layout that is a contiguous slab of rows -- together with the
range rows attached to them, and multiplies only its rows:

  * every measurement with an owned pose at either end is assembled into the rank's LOCAL data matrix, so the rows of
    the owned poses and ranges are complete; the poses at the far end of such measurements are GHOSTS (for an
    odometry chain: one pose either side of the slab), whose rows are incomplete and never used;
  * the landmark rows are replicated: every rank holds all landmark values, computes the partial sums of the landmark
    rows over ITS measurements, and the partials are all-reduced (l x r doubles);
  * before a product the ghost rows of the operand are fetched from their owners, GPU to GPU over NVLink; the local
    product is the library's persistent SpMM kernel on the local handle.  Two formulations of the exchange:
    `RowPartitionedProduct` (collectives through `torch.distributed`: one all-to-all of the ghost pose blocks, one
    all-reduce of the landmark rows) and `peer_product` (the library's own kernels over peer-mapped memory,
    cora_b200/csrc/peer_product.cuh: flag barrier + pull, rank-ordered sum; no collective per product).

No arithmetic happens here: this module builds index sets (host) and moves rows (`index_select` / `index_copy_` on the
device buffers the C-ABI exposes, `all_to_all_single`, `all_reduce`).  `product_fn` abstracts the local product so the
partition and exchange logic is covered on CPU (`gloo`, tests/test_rowpart_cpu.py) with a SciPy stand-in.

Measured at 1M poses, rank 5 (profiles/r02_rowpart.jsonl): 161 us per product on one GPU; with collectives 211 / 174 /
174 us on 2 / 4 / 8 GPUs (two ~35 us collectives eat the gain, as SURVEY 8(e) predicts); with the peer-memory kernels
116 / 82 / 64 us.  Replicas remain the throughput mode at the BASELINE sizes (bench.py).
"""
import numpy as np


class LocalProblem:
    """The slab of rank `rank` out of `world`: local measurement arrays + maps between local and global rows
    (reference row order [d n | m | n | l] on both sides)."""

    def __init__(self, d, n, l, arrays, world, rank):
        self.d, self.n, self.l, self.world, self.rank = int(d), int(n), int(l), int(world), int(rank)
        A = {k: np.asarray(v) for k, v in arrays.items()}
        bounds = (np.arange(world + 1, dtype=np.int64) * n) // world
        self.bounds = bounds
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        self.lo, self.hi = lo, hi
        owner_of_pose = lambda p: np.searchsorted(bounds, p, side="right") - 1
        own = lambda idx: (idx >= lo) & (idx < hi)            # pose index (or translation index of a pose) owned here
        is_pose = lambda idx: idx < n
        # relative-translation factors (first id is a pose; second a pose or a landmark) and rotation factors
        rp_keep = own(A["rp_i"]) | (is_pose(A["rp_j"]) & own(A["rp_j"]))
        rot_keep = own(A["rot_i"]) | own(A["rot_j"])
        a, b = A["rg_a"], A["rg_b"]
        pa, pb = is_pose(a), is_pose(b)
        if np.any(pa & pb & (owner_of_pose(np.where(pa, a, 0)) != owner_of_pose(np.where(pb, b, 0)))):
            raise NotImplementedError("a range between poses of different slabs needs a ghost range row")
        # a range row belongs to the slab of its pose (landmark-landmark ranges: rank 0)
        rg_owner = np.where(pa, owner_of_pose(np.where(pa, a, 0)), np.where(pb, owner_of_pose(np.where(pb, b, 0)), 0))
        rg_keep = rg_owner == rank
        self.rg_global = np.nonzero(rg_keep)[0]                # global range ids of the local range rows, in order
        # poses of the local problem: owned first, ghosts (far ends of kept factors) after, by global id
        ends = np.concatenate([A["rp_i"][rp_keep], A["rp_j"][rp_keep][is_pose(A["rp_j"][rp_keep])],
                               A["rot_i"][rot_keep], A["rot_j"][rot_keep]]).astype(np.int64)
        ghosts = np.setdiff1d(np.unique(ends), np.arange(lo, hi))
        self.poses = np.concatenate([np.arange(lo, hi, dtype=np.int64), ghosts])   # local pose id -> global pose id
        self.n_own, self.n_loc = hi - lo, len(self.poses)
        loc_of = np.full(n, -1, dtype=np.int64)
        loc_of[self.poses] = np.arange(self.n_loc)
        tr = lambda idx: np.where(idx < n, loc_of[np.minimum(idx, n - 1)], self.n_loc + (idx - n))  # translation index
        if np.any(tr(a[rg_keep]) < 0) or np.any(tr(b[rg_keep]) < 0):
            raise NotImplementedError("a local range touches a pose that is neither owned nor a ghost")
        self.arrays = dict(
            rp_i=tr(A["rp_i"][rp_keep]), rp_j=tr(A["rp_j"][rp_keep]), rp_t=A["rp_t"][rp_keep], rp_tau=A["rp_tau"][rp_keep],
            rot_i=loc_of[A["rot_i"][rot_keep]], rot_j=loc_of[A["rot_j"][rot_keep]], rot_R=A["rot_R"][rot_keep],
            rot_kappa=A["rot_kappa"][rot_keep], rg_a=tr(a[rg_keep]), rg_b=tr(b[rg_keep]), rg_r=A["rg_r"][rg_keep],
            rg_w=A["rg_w"][rg_keep])
        self.m_loc, self.m = int(rg_keep.sum()), len(a)
        self.N_loc = (d + 1) * self.n_loc + self.m_loc + l
        self.N = (d + 1) * n + self.m + l
        # local reference row -> global reference row
        g = np.empty(self.N_loc, dtype=np.int64)
        dn_l, dn_g = d * self.n_loc, d * n
        g[:dn_l] = (self.poses[:, None] * d + np.arange(d)[None, :]).reshape(-1)
        g[dn_l: dn_l + self.m_loc] = dn_g + self.rg_global
        g[dn_l + self.m_loc: dn_l + self.m_loc + self.n_loc] = dn_g + self.m + self.poses
        g[dn_l + self.m_loc + self.n_loc:] = dn_g + self.m + n + np.arange(l)
        self.local_to_global = g
        # which local rows this rank owns (landmark rows: replicated, reduced over the ranks)
        owned = np.zeros(self.N_loc, dtype=bool)
        owned[: d * self.n_own] = True
        owned[dn_l: dn_l + self.m_loc] = True
        owned[dn_l + self.m_loc: dn_l + self.m_loc + self.n_own] = True
        self.owned = owned
        self.landmark_rows = np.arange(dn_l + self.m_loc + self.n_loc, self.N_loc)
        self.ghost_owner = owner_of_pose(ghosts) if len(ghosts) else np.zeros(0, dtype=np.int64)

    def pose_rows(self, local_pose_ids):
        """Local reference rows (d rotation rows + the translation row) of the given local poses, pose by pose."""
        d = self.d
        p = np.asarray(local_pose_ids, dtype=np.int64)
        rot = (p[:, None] * d + np.arange(d)[None, :])
        trn = (d * self.n_loc + self.m_loc + p)[:, None]
        return np.concatenate([rot, trn], axis=1).reshape(-1)


def exchange_plan(parts):
    """For every rank: which of its OWNED poses each peer needs as a ghost (send lists) and where its own ghosts come
    from (receive lists).  `parts`: the LocalProblem of every rank (cheap: index arrays only; each rank can build all
    of them from the global arrays, or only its own and exchange the ghost lists)."""
    world = len(parts)
    send = [[np.zeros(0, dtype=np.int64) for _ in range(world)] for _ in range(world)]  # send[src][dst]: global pose ids
    for dst, P in enumerate(parts):
        ghosts = P.poses[P.n_own:]
        for src in range(world):
            send[src][dst] = ghosts[P.ghost_owner == src]
    return send


class RowPartitionedProduct:
    """Y = Q X with X, Y distributed by rows.  `product_fn()` multiplies the LOCAL operand buffer into the local result
    buffer (the persistent SpMM kernel of the rank's handle, or a SciPy stand-in on CPU); `x_buf` / `y_buf` are torch
    tensors (N_loc x r) over those buffers in the row order `row_of_local_ref` (local reference row -> buffer row)."""

    def __init__(self, part, send_lists, x_buf, y_buf, row_of_local_ref, product_fn, dist, group=None):
        import torch
        self.P, self.dist, self.group, self.torch = part, dist, group, torch
        self.x, self.y, self.product_fn = x_buf, y_buf, product_fn
        dev = x_buf.device
        row_of = np.asarray(row_of_local_ref, dtype=np.int64)
        glob2loc = {int(g): i for i, g in enumerate(part.poses)}
        me, world = part.rank, part.world
        D1 = part.d + 1
        to_rows = lambda poses: row_of[part.pose_rows([glob2loc[int(p)] for p in poses])] if len(poses) else np.zeros(0, np.int64)
        self.send_rows = torch.as_tensor(np.concatenate([to_rows(send_lists[me][q]) for q in range(world)]), device=dev)
        self.send_split = [len(send_lists[me][q]) * D1 for q in range(world)]
        self.recv_rows = torch.as_tensor(np.concatenate([to_rows(send_lists[q][me]) for q in range(world)]), device=dev)
        self.recv_split = [len(send_lists[q][me]) * D1 for q in range(world)]
        self.lm_rows = torch.as_tensor(row_of[part.landmark_rows], device=dev)
        r = x_buf.shape[1]
        self.send_buf = torch.empty((int(sum(self.send_split)), r), dtype=x_buf.dtype, device=dev)
        self.recv_buf = torch.empty((int(sum(self.recv_split)), r), dtype=x_buf.dtype, device=dev)
        self.lm_buf = torch.empty((len(part.landmark_rows), r), dtype=x_buf.dtype, device=dev)

    def halo(self):
        """Ghost rows of the operand <- their owners."""
        t = self.torch
        if self.P.world == 1:
            return
        t.index_select(self.x, 0, self.send_rows, out=self.send_buf)
        self.dist.all_to_all_single(self.recv_buf, self.send_buf, self.recv_split, self.send_split, group=self.group)
        self.x.index_copy_(0, self.recv_rows, self.recv_buf)

    def reduce_landmarks(self):
        t = self.torch
        if self.P.world == 1 or len(self.P.landmark_rows) == 0:
            return
        t.index_select(self.y, 0, self.lm_rows, out=self.lm_buf)
        self.dist.all_reduce(self.lm_buf, group=self.group)
        self.y.index_copy_(0, self.lm_rows, self.lm_buf)

    def __call__(self):
        self.halo()
        self.product_fn()
        self.reduce_landmarks()


class _DevArray:
    """`__cuda_array_interface__` view of a raw device pointer (N x r f64, row-major) for torch.as_tensor."""

    def __init__(self, ptr, rows, cols):
        self.__cuda_array_interface__ = {"shape": (int(rows), int(cols)), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_product(handle, part, send_lists, r, dist, group=None):
    """RowPartitionedProduct over the device buffers of `handle` (the local problem's capi.Handle)."""
    import torch
    xp, yp = handle.device_vectors(r)
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.as_tensor(_DevArray(xp, part.N_loc, r), device=dev)
    y = torch.as_tensor(_DevArray(yp, part.N_loc, r), device=dev)
    int2ref = handle.row_order().astype(np.int64)
    row_of = np.empty(part.N_loc, dtype=np.int64)
    row_of[int2ref] = np.arange(part.N_loc)
    stream_sync = torch.cuda.current_stream().synchronize

    def product():
        stream_sync()              # the exchange ran on torch's stream, the kernel runs on the handle's
        handle.spmm_resident(1)    # (returns after its own CUDA events: the result is complete)

    return RowPartitionedProduct(part, send_lists, x, y, row_of, product, dist, group), x, y, row_of


def internal_rows_of_pose(d, local_pose_id):
    """Rows of a pose in the library's internal (pose-major) layout: d rotation rows, then the translation row
    (cora_b200/csrc/layout.hpp); checked against cora_b200_row_order by peer_product()."""
    return (d + 1) * int(local_pose_id) + np.arange(d + 1)


def peer_product(handle, parts, rank, r, dist, group=None):
    """The same product with the library's own exchange kernels over peer-mapped memory (cora_b200_peer_*,
    cora_b200/csrc/peer_product.cuh): no collective per product.  `dist` only ships the 192-byte IPC handles once."""
    from . import capi
    P = parts[rank]
    world, d = P.world, P.d
    int2ref = handle.row_order().astype(np.int64)
    row_of = np.empty(P.N_loc, dtype=np.int64)
    row_of[int2ref] = np.arange(P.N_loc)
    # the pose-major formula every rank uses for its PEERS' buffers must hold for this rank's own buffer
    probe = [0, P.n_own - 1, P.n_loc - 1]
    for p in probe:
        assert np.array_equal(row_of[P.pose_rows([p])], internal_rows_of_pose(d, p)), "unexpected internal row order"
    assert np.array_equal(row_of[P.landmark_rows], (d + 1) * P.n_loc + np.arange(P.l)), "unexpected landmark rows"
    ghosts = P.poses[P.n_own:]
    gp, gs, gd = [], [], []
    for k, g in enumerate(ghosts):
        q = int(P.ghost_owner[k])
        gp += [q] * (d + 1)
        gs += list(internal_rows_of_pose(d, int(g) - int(P.bounds[q])))   # owned poses come first on their owner
        gd += list(internal_rows_of_pose(d, P.n_own + k))
    pp = capi.PeerProduct(handle, r, P.l)
    if world > 1:
        allh = [None] * world
        dist.all_gather_object(allh, pp.handles, group=group)
        allh = b"".join(allh)
    else:
        allh = pp.handles
    pp.connect(world, rank, allh, gp, gs, gd, row_of[P.landmark_rows])
    return pp, row_of
