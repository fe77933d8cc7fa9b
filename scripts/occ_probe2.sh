#!/bin/bash
# per-SM throughput of register-capped builds of the chain kernel at N resident warps per SM (SRHMC_CHAIN_MAX_RESIDENT)
# usage: scripts/occ_probe2.sh tag:cap[,cap...] ...
for spec in "$@"; do
tag=${spec%%:*}; caps=${spec#*:}
lib=/root/repo/hmc_stellar_toy_model_b200/libstellar_rhmc_$tag.so
for k in ${caps//,/ }; do
  SRHMC_LIB=$lib SRHMC_CHAIN_MAX_RESIDENT=$k timeout 300 python bench.py --workload c2 --steps 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag warps/SM <= $k', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms')"
done
done
