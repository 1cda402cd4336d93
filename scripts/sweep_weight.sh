#!/bin/bash
# sweep of the scalar-row tile weight of the CTA partition (CORA_B200_SCALAR_TILE_WEIGHT), bench workload
for w in "$@"; do
  CORA_B200_SCALAR_TILE_WEIGHT=$w timeout 300 python bench.py --no-cpu-baseline --no-solve --steps 5 > gpurun_out/sw_$w.log 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/sw_$w.log").read().strip().splitlines()[-1]); p=d["roofline"]["phases_in_kernel_globaltimer_cta0"]
print("w=$w  us/CG %.1f  hess %.1f update %.1f pupdate %.1f hub %.1f sync %.2f" % (d["us_per_cg_iteration"], p["hess"]["avg_us"], p["update"]["avg_us"], p["pupdate"]["avg_us"], p["hub"]["avg_us"], p["sync"]["avg_us"]))
PY
done
