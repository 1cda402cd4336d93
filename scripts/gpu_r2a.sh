#!/bin/bash
# round 2, first check of the streaming kernels: parity tests, bench, per-CTA phase profile, 1M-pose sweep
NAME=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_tnt.py tests/test_gpu_operators.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --no-cpu-baseline --no-solve > gpurun_out/${NAME}_bench.log 2>gpurun_out/${NAME}_bench.err; tail -n 5 gpurun_out/${NAME}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${NAME}_bench.log").read().strip().splitlines()[-1])
print("value %.1f  us/CG %.1f  e2e %.1f  frac %.3f" % (d["value"],d["us_per_cg_iteration"],d["e2e"]["value"],d["roofline"]["frac"]))
for k,v in d["roofline"]["phases_in_kernel_globaltimer_cta0"].items(): print("  %-8s %8.2f us x %d" % (k, v["avg_us"], v["count"]))
PY
CORA_B200_STREAM=0 timeout 300 python bench.py --no-cpu-baseline --no-solve 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('TILE PATH value %.1f  us/CG %.1f' % (d['value'],d['us_per_cg_iteration']))"
CORA_B200_PHASE_PROFILE=1 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | tail -40
timeout 600 python scripts/sweep_1m.py 1000000 30 2 2>&1 | tail -5
