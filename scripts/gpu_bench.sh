#!/bin/bash
# full GPU suite + the complete default bench line (both arms)
NAME=${1:-b}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${NAME}_bench.log 2>gpurun_out/${NAME}_bench.err; tail -n 5 gpurun_out/${NAME}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${NAME}_bench_ref.log 2>gpurun_out/${NAME}_bench_ref.err; tail -n 1 gpurun_out/${NAME}_bench_ref.log | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/${NAME}_bench.log").read().strip().splitlines()[-1])
print("value %.1f  us/CG %.1f  e2e %.1f  frac %.3f cpu %s" % (d["value"],d["us_per_cg_iteration"],d["e2e"]["value"],d["roofline"]["frac"],d.get("cpu_baseline",{}).get("value")))
print("  ".join("%s %.2f/%.2f" % (k, v["avg_us"], v.get("slowest_cta_avg_us",0)) for k,v in d["roofline"]["phases_in_kernel_globaltimer"].items()))
print("cfg5", json.dumps(d["roofline_cfg5"])[:900])
print("cfg2", json.dumps(d["spmv_cfg2"]))
print("solve", json.dumps(d["solve_to_cert"])[:1500])
PY
