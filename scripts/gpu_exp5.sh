#!/bin/bash
for n in 100000 1000000; do
echo "=== n=$n"
CORA_B200_LIB=$PWD/cora_b200/lib/exp_SUBPROF.so CORA_B200_PHASE_PROFILE=1 timeout 600 python scripts/profile_cg.py 2 $n 1 3 2>&1 | grep -A45 "outer 2" | grep "per-CTA avg q\|per-CTA avg ch\|per-CTA avg hess"
done
