#!/bin/bash
# end-of-round evidence: tests, smoke, bench (both arms), ncu launch list, ncu --set full of the persistent kernels
NAME=${1:-r2z}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${NAME}_bench.log 2>gpurun_out/${NAME}_bench.err; tail -n 3 gpurun_out/${NAME}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${NAME}_bench_ref.log 2>gpurun_out/${NAME}_bench_ref.err; tail -n 1 gpurun_out/${NAME}_bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${NAME}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${NAME}_bench_under_ncu.log 2>&1
timeout 900 bash scripts/profile_persistent.sh 2 12 ${NAME}_persistent
ncu -i gpurun_out/${NAME}_persistent.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
timeout 600 bash scripts/gpu_ncu_spmm.sh ${NAME}_spmm1m 1000000
ncu -i gpurun_out/${NAME}_spmm1m.ncu-rep --page raw --csv > gpurun_out/${NAME}_spmm1m_raw.csv 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/${NAME}_bench.log").read().strip().splitlines()[-1])
print("value %.1f  us/CG %.1f  e2e %.1f  frac %.3f cpu %.1f" % (d["value"],d["us_per_cg_iteration"],d["e2e"]["value"],d["roofline"]["frac"],d["cpu_baseline"]["value"]))
for k,v in d["roofline"]["phases_in_kernel_globaltimer"].items(): print("  %-8s %8.2f us x %d  (slowest CTA %.2f)" % (k, v["avg_us"], v["count"], v.get("slowest_cta_avg_us", 0)))
PY
