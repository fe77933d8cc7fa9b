#!/bin/bash
# ncu evidence for the large-field engine (run under gpurun, one GPU).  Usage: scripts/profile_c5.sh <tag> [bench args]
set -u
TAG=${1:-r1}; shift
CMD="python bench.py --workload c5 --steps 2 --warmup 3 --e2e-steps 1 --niter 2 $*"
$CMD > gpurun_out/c5_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 160 --csv --log-file gpurun_out/c5_launches_$TAG.csv $CMD > gpurun_out/c5_ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/c5_plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-big_tile} -s 40 -c 1 -o gpurun_out/c5_prof_$TAG $CMD > gpurun_out/c5_ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/c5_plain_$TAG.log | cut -c1-300
