#!/bin/bash
# ncu --set full of the heaviest kernels (one launch each) on a 2048-service hotel stream
set -u
T=${1:-a}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_score3 -s 6 -c 2 -o gpurun_out/${T}_score3 -f \
    python scripts/profile_run.py 2048 > gpurun_out/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stitch -s 3 -c 1 -o gpurun_out/${T}_stitch -f \
    python scripts/profile_run.py 2048 >> gpurun_out/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gmm_bic -s 9 -c 1 -o gpurun_out/${T}_gmm_bic5 -f \
    python scripts/profile_run.py 2048 >> gpurun_out/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gmm_lloyd -s 18 -c 1 -o gpurun_out/${T}_gmm_lloyd5 -f \
    python scripts/profile_run.py 2048 >> gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -5 gpurun_out/${T}_ncu_full.log
