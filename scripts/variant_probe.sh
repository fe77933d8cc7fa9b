#!/bin/bash
# C2 headline value under each experimental library build given on the command line (tags of libstellar_rhmc_<tag>.so; "base" = the shipped one)
for tag in "$@"; do
  lib=/root/repo/hmc_stellar_toy_model_b200/libstellar_rhmc_$tag.so
  [ "$tag" = base ] && lib=/root/repo/hmc_stellar_toy_model_b200/libstellar_rhmc.so
  SRHMC_LIB=$lib timeout 300 python bench.py --workload c2 --steps 4 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms')"
done
