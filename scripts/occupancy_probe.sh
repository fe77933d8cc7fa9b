#!/bin/bash
# Per-round time versus resident warps per SM: 6512 chains = 1628 warps = 11 per SM, run at k = 11, 6, 4, 3, 2, 1.
for k in 11 6 4 3 2 1; do
  SRHMC_CHAIN_CHUNKS=1 SRHMC_CHAIN_BLOCKS_PER_SM=$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --chains-per-mag 592 --niter 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('k=$k value %.1f M/s  ms %.2f' % (d['value']/1e6, d['ms_per_step']))"
done
