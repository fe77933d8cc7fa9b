#!/bin/bash
# Usage: [SRHMC_CHAIN_CHUNKS=c] scripts/sweep_chain.sh <k> [<k> ...]
# headline workload at k resident warps per SM
for k in "$@"; do
  SRHMC_CHAIN_BLOCKS_PER_SM=$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks=${SRHMC_CHAIN_CHUNKS:-auto} k=$k value %.1f M/s  ms %.1f frac %.3f acc %.4f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['accept_rate']))"
done
