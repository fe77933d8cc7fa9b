#!/bin/bash
# Usage: scripts/sweep_chain.sh <lpc> <k> [<k> ...]   -- headline workload at the given warps-per-SM settings
LPC=$1; shift
for k in "$@"; do
  SRHMC_CHAIN_LPC=$LPC SRHMC_CHAIN_BLOCKS_PER_SM=$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('LPC$LPC k=$k value %.1f M/s  ms %.1f frac %.3f acc %.4f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['accept_rate']))"
done
