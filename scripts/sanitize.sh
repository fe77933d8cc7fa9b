#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (run under gpurun, one GPU):
#   scripts/sanitize.sh [tag]   ->  gpurun_out/sanitize_<tag>_{memcheck,racecheck,synccheck}.txt (copy the summaries to profiles/)
# racecheck covers the shared-memory hazards of the barrier-free warp-private layout of the chain kernel and of the tile /
# field kernels; memcheck the global accesses incl. the peer mailboxes; synccheck the barrier / mbarrier usage.
set -u
TAG=${1:-r2}
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck; do
  for part in chain field big; do
    out=gpurun_out/sanitize_${TAG}_${tool}_${part}.txt
    timeout 900 $SAN --tool $tool --print-limit 20 python scripts/sanitize_workload.py $part > $out 2>&1
    echo "$tool $part rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out | tail -1)"
  done
done
