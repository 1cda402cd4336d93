#!/bin/bash
NAME=${1:-q}
CORA_B200_PHASE_PROFILE=1 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | grep -A40 "CG 240" | grep "n=\|per-CTA avg q\|per-CTA avg hess"
