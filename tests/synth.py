"""Synthetic problems for the tests: product-side generator -> oracle Problem."""
from cora_b200 import synthetic
from oracle import cora_oracle as co


def make_synthetic(n, l, m, d=3, seed=42, rank=None, preconditioner=co.JACOBI, loop_closures=None):
    arrays, gt = synthetic.make_arrays(n, l, m, d=d, seed=seed, loop_closures=loop_closures)
    p = co.Problem.from_arrays(d, n, l, arrays, rank=rank, preconditioner=preconditioner)
    p.gt = gt
    return p
