#!/bin/bash
# quick loop: streaming-kernel parity tests + bench line + phase profile + 1M SpMM
NAME=${1:-q}
timeout 900 python -m pytest tests/test_gpu_tnt.py tests/test_gpu_operators.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-solve 2>/dev/null > gpurun_out/${NAME}_bench.log
python - <<PY
import json
d=json.loads(open("gpurun_out/${NAME}_bench.log").read().strip().splitlines()[-1])
print("BENCH value %.1f  us/CG %.1f  e2e %.1f  frac %.3f" % (d["value"],d["us_per_cg_iteration"],d["e2e"]["value"],d["roofline"]["frac"]))
print("  ".join("%s %.2f" % (k, v["avg_us"]) for k,v in d["roofline"]["phases_in_kernel_globaltimer"].items()))
PY
CORA_B200_PHASE_PROFILE=1 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | grep "per-CTA" | tail -10
timeout 300 python scripts/profile_cg.py 30 1000000 1 12 spmm 2>&1 | tail -1
