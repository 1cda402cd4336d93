"""Generate the committed golden fixtures from the reference's own test data.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Outputs (committed, small):
  tests/golden/<fixture>.npz  for the three Catch2 fixture problems of
      /root/reference/tests/data/{small_ra_slam_problem,single_rpm,single_range}
      (factor_graph.pyfg text + the 16 MatrixMarket goldens, densified) and the
      expected costs of tests/test_utils.cpp:213-217;
  tests/golden/{plaza2,single_drone,tiers,mrclam2}.npz  measurement arrays of the two real
      datasets BASELINE.json names (examples/data/*.pyfg) flattened with the
      oracle's parser, so the GPU box (which has no /root/reference) can rebuild
      the exact same data matrix.

Nothing in tests/, bench.py or smoke() reads /root/reference at run time.
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import cora_oracle as co  # noqa: E402

REF = "/root/reference"
MM = ["Apose", "Arange", "OmegaPose", "OmegaRange", "RangeDistances", "T", "RotConLaplacian",
      "DataMatrix", "S_rand", "X_gt", "X_odom", "X_rand_dim2", "expected_egrad",
      "expected_rgrad", "hessProd", "rand_dX"]
COST = {"small_ra_slam_problem": 1.063888372855624e+03, "single_rpm": 0.809173848024762,
        "single_range": 4.718031199983851}  # tests/test_utils.cpp:213-217


def dense(path):
    M = sio.mmread(path)
    return np.asarray(M.todense()) if hasattr(M, "todense") else np.asarray(M)


def main():
    for name, cost in COST.items():
        src = os.path.join(REF, "tests", "data", name)
        out = {"pyfg": np.array(open(os.path.join(src, "factor_graph.pyfg")).read()),
               "expected_cost": np.array(cost)}
        for k in MM:
            out[k] = dense(os.path.join(src, k + ".mm"))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k not in ("pyfg",)})
    # tiers / mrclam2: several robots + inter-robot measurements (general sparse Cholesky)
    files = {"plaza2": "plaza2.pyfg", "single_drone": "single_drone.pyfg", "tiers": "tiers.pyfg",
             "mrclam2": os.path.join("mrclam", "range_and_rpm", "mrclam2", "mrclam2.pyfg")}
    only = sys.argv[1:]
    for name, rel in files.items():
        if only and name not in only:
            continue
        p = co.parse_pyfg(os.path.join(REF, "examples", "data", rel))
        a = p.measurement_arrays()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), d=p.d, n=p.n, l=p.l, **a)
        print(name, p.d, p.n, p.l, p.m, len(a["rp_tau"]))


if __name__ == "__main__":
    main()
