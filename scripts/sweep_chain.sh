#!/bin/bash
# sweep resident warps/SM of the chain kernel on the headline workload
for k in "$@"; do
  SRHMC_CHAIN_BLOCKS_PER_SM=$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('k=$k value %.1f M/s  ms %.1f frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['frac']))"
done
