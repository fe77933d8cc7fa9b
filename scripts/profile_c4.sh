#!/bin/bash
# ncu evidence for the crowded-field kernel (run under gpurun, one GPU).  Usage: scripts/profile_c4.sh <tag>
set -u
TAG=${1:-c4}
CMD="python bench.py --workload c4 --fields 296 --niter 3 --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:field_kernel -s 3 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
