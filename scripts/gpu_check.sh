#!/bin/bash
# tests + smoke + bench summary on the GPU box
NAME=${1:-check}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${NAME}_bench.log 2>gpurun_out/${NAME}_bench.err; tail -n 5 gpurun_out/${NAME}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${NAME}_bench.log").read().strip().splitlines()[-1])
print("value %.1f  us/CG %.1f  e2e %.1f  frac %.3f" % (d["value"],d["us_per_cg_iteration"],d["e2e"]["value"],d["roofline"]["frac"]))
for k,v in d["roofline"]["phases_in_kernel_globaltimer"].items(): print("  %-8s %8.2f us x %d" % (k, v["avg_us"], v["count"]))
PY
