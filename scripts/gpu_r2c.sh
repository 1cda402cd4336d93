#!/bin/bash
NAME=${1:-r2c}
timeout 900 python -m pytest tests/test_gpu_tnt.py tests/test_gpu_operators.py -m gpu -x -q 2>&1 | tail -5
CORA_B200_PHASE_PROFILE=2 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | tail -75 > gpurun_out/${NAME}_percta.log
grep -A28 "per-CTA avg hess" gpurun_out/${NAME}_percta.log | tail -29
tail -22 gpurun_out/${NAME}_percta.log | grep -v "@"
timeout 300 python bench.py --no-cpu-baseline --no-solve 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH value %.1f  us/CG %.1f e2e %.1f frac %.3f' % (d['value'],d['us_per_cg_iteration'],d['e2e']['value'],d['roofline']['frac']))"
timeout 300 python scripts/profile_cg.py 30 1000000 1 12 spmm 2>&1 | tail -1
