#!/bin/bash
NAME=${1:-spmm}; N=${2:-100000}
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spmm_persistent -c 1 \
    -f -o gpurun_out/${NAME} python scripts/profile_cg.py 8 $N 1 2 spmm > gpurun_out/${NAME}.log 2>&1
tail -n 2 gpurun_out/${NAME}.log
