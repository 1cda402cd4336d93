#!/bin/bash
# ncu --set full capture of the persistent TNT kernel on the bench workload (one launch = `outer` TNT iterations)
OUTER=${1:-2}; PRE=${2:-12}; NAME=${3:-persistent}
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_tnt_persistent -c 1 \
    -f -o gpurun_out/${NAME} python scripts/profile_cg.py $OUTER 100000 1 $PRE > gpurun_out/${NAME}.log 2>&1
tail -n 3 gpurun_out/${NAME}.log
ls -la gpurun_out/${NAME}.ncu-rep
