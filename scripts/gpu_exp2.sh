#!/bin/bash
# A/B of runtime knobs: each argument is an "ENV=VAL,ENV=VAL" set
for cfg in "$@"; do
  echo "=== $cfg"
  envs=$(echo $cfg | tr ',' ' ')
  env $envs CORA_B200_PHASE_PROFILE=1 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | grep -A30 "CG 240" | grep "n=.*hess\|n=.*update\|q.wait\|q.qx  \|device\|per-CTA avg hess" | head -8
  env $envs timeout 300 python scripts/profile_cg.py 30 1000000 1 12 spmm 2>&1 | tail -1
done
