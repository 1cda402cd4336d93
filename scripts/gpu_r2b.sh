#!/bin/bash
NAME=${1:-r2b}
CORA_B200_PHASE_PROFILE=2 timeout 300 python scripts/profile_cg.py 3 100000 1 12 2>&1 | tail -75 > gpurun_out/${NAME}_percta.log
tail -70 gpurun_out/${NAME}_percta.log
timeout 900 bash scripts/profile_persistent.sh 2 12 ${NAME}_persistent
ncu -i gpurun_out/${NAME}_persistent.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
timeout 300 python scripts/profile_cg.py 30 1000000 1 12 spmm 2>&1 | tail -2
