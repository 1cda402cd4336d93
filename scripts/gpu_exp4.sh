#!/bin/bash
# SpMM-only timing of alternative builds at 100k (L2 resident) and 1M poses
for v in "$@"; do
  echo "=== $v"
  CORA_B200_LIB=$PWD/cora_b200/lib/$v timeout 300 python scripts/profile_cg.py 200 100000 1 2 spmm 2>&1 | tail -1
  CORA_B200_LIB=$PWD/cora_b200/lib/$v timeout 300 python scripts/profile_cg.py 30 1000000 1 2 spmm 2>&1 | tail -1
done
