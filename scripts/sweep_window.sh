timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
P='import json,sys; d=json.loads(sys.stdin.read()); print("%s value %.1f M/s  ms %.1f frac %.3f acc %.4f" % (sys.argv[1], d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["accept_rate"]))'
SRHMC_CHAIN_WINDOW=0 $B 2>/dev/null | python -c "$P" full
$B 2>/dev/null | python -c "$P" win
for v in r200 u3 u1; do SRHMC_LIB=$PWD/hmc_stellar_toy_model_b200/libstellar_rhmc_$v.so $B 2>/dev/null | python -c "$P" $v; done
for k in 8 9; do SRHMC_CHAIN_BLOCKS_PER_SM=$k $B 2>/dev/null | python -c "$P" win_k$k; done
