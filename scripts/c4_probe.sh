#!/bin/bash
# Crowded-field (C4) throughput of the CTA-per-field kernel under launch-shape experiments (env is inherited).
python bench.py --workload c4 --fields ${FIELDS:-592} --niter 10 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --precision ${PREC:-64} 2>&1 | tail -1 | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print('c4 thr=%s smem=%s dglob=%s prec=%s: %.1f M/s ms %.1f frac %.3f' % (os.environ.get('SRHMC_FIELD_THREADS','def'), os.environ.get('SRHMC_FIELD_SMEM_KB','def'), os.environ.get('SRHMC_FIELD_D_GLOBAL','0'), os.environ.get('PREC','64'), d['value']/1e6, d['ms_per_step'], d['roofline']['frac']))"
