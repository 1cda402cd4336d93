#!/bin/bash
for v in "$@"; do
  echo "=== $v"
  CORA_B200_LIB=$PWD/cora_b200/lib/$v CORA_B200_PHASE_PROFILE=1 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | grep -A45 "CG 240" | grep "n=\|per-CTA avg q\|per-CTA avg ch\|per-CTA avg hess"
done
