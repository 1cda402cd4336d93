#!/bin/bash
# timing-only experiments: alternative builds selected through CORA_B200_LIB
for v in "$@"; do
  echo "=== $v"
  CORA_B200_LIB=$PWD/cora_b200/lib/$v CORA_B200_PHASE_PROFILE=1 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | grep -A30 "CG 240" | grep "hess\|update \|q.qx\|device" | head -8
  CORA_B200_LIB=$PWD/cora_b200/lib/$v timeout 300 python scripts/profile_cg.py 30 1000000 1 12 spmm 2>&1 | tail -1
done
