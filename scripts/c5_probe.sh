#!/bin/bash
# C5 single-GPU probe: value (M star-steps/s) of `bench.py --workload c5` under the environment given on the command line
# usage: scripts/c5_probe.sh TAG [VAR=value ...]
TAG=$1; shift
env "$@" python bench.py --workload c5 --steps 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$TAG', '$*', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms', round(d['roofline']['frac'],3))"
