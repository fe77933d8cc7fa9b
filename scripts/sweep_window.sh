#!/bin/bash
# Headline workload with the one-star kernel's column window on / off and at other row cuts (run under gpurun).
B="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
P='import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print("%s value %.1f M/s  ms %.1f frac %.3f acc %.4f" % (sys.argv[1], d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["accept_rate"]))'
SRHMC_CHAIN_WINDOW=0 $B 2>/dev/null | python -c "$P" full_width
$B 2>/dev/null | python -c "$P" window24
SRHMC_CHAIN_WCUT_BITS=50 $B 2>/dev/null | python -c "$P" window24_rowcut50
