B="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
P='import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print("%s %.1f M/s ms %.2f" % (sys.argv[1], d["value"]/1e6, d["ms_per_step"]))'
for c in 2 6 10 16 24 40; do SRHMC_CHAIN_CHUNKS=$c $B 2>/dev/null | python -c "$P" chunks$c; done
