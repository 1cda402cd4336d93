#!/bin/bash
CORA_B200_PHASE_PROFILE=1 timeout 600 python scripts/profile_cg.py 2 1000000 1 3 2>&1 | grep -A45 "outer 2" | grep "n=\|per-CTA avg q\|per-CTA avg hess\|per-CTA avg update\|per-CTA avg sync"
